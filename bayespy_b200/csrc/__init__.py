"""CUDA sources of libbpk.so (sm_100a only) and the in-tree nvcc build script (build.py)."""
