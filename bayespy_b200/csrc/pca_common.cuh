// pca_common.cuh — shapes shared by the factor-model kernels (pca.cu, pca_vb.cu).
#pragma once
#define PCA_MP 64      // padded M of the tensor-pipe sweep kernel
#define PCA_KP 16      // padded K
// per-CTA partial statistics written by pca_xsweep_kernel: [S_yx 64x16 | S_xx 16x16 | s_x 16]
#define PCA_NSTAT (PCA_MP * PCA_KP + PCA_KP * PCA_KP + PCA_KP)

