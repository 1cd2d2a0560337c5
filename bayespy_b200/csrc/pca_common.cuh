// pca_common.cuh — shapes shared by the factor-model kernels (pca.cu, pca_vb.cu).
#pragma once
#define PCA_MP 64      // padded M of the tensor-pipe sweep kernel
#define PCA_KP 16      // padded K
// per-CTA partial statistics written by pca_xsweep_kernel: [S_yx 64x16 | S_xx 16x16 | s_x 16]
#define PCA_NSTAT (PCA_MP * PCA_KP + PCA_KP * PCA_KP + PCA_KP)

// pca.cu: the one-pass sweep with the grid reduction left to the caller (requires M<=64, K<=16)
int pca_xsweep_partials(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                        double *X, const int *stop, double **partial_out, int *nparts_out);
