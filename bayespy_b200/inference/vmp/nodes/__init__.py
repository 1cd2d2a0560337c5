"""Import paths of the reference for the few internals model scripts use (bayespy/inference/vmp/nodes/)."""
