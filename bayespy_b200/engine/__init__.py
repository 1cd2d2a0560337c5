"""Host-side node graph of the engine: plates, masks, moments and messages on device arrays."""
