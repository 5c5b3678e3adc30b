// The orthogonal projector onto range(F), P = F pinv(F), on the device in fp64
// (ProjectionMatrixBuilder.get_projection_matrix, m_diffuser/dynamics/projection.py:85-120: numpy's SVD-based pinv,
// cast to fp32 at the end).  F = [state rollout rows; I] always has full column rank -- its rows contain the
// identity on (x0, u_0..u_{H-1}) -- so P = F (F^T F)^{-1} F^T = W^T W with W = L^{-1} F^T and F^T F = L L^T:
//   1. G = F^T F                                   (dgemm)
//   2. G = L L^T, blocked right-looking Cholesky   (32-wide panels: diagonal block, panel solve, dgemm update)
//   3. W = L^{-1} F^T, blocked forward substitution (diagonal solve, dgemm update)
//   4. P = W^T W                                   (dgemm), cast to fp32
// ~15 GFLOP for Door (D = 2183, rank 935): 15 ms here including both host copies, 280 ms in numpy on the GPU box's
// host -- the dynamics or the horizon can change online.  Every kernel is plain fp64 SIMT; this is a once-per-(A, B, H) precompute, not the hot path.
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace dad {

// C[i, j] = beta C[i, j] + alpha sum_k A(i, k) B(k, j), all operands addressed by element strides so that any
// transpose / sub-matrix view is one launch.  64x64 tiles, K step 16, 256 threads, 4x4 outputs per thread.
struct GemmD {
  const double *A, *B;
  double *C;
  int M, N, K;
  long long a_i, a_k, b_k, b_j, c_i, c_j;
  double alpha, beta;
};

__global__ void __launch_bounds__(256) dgemm_kernel(const GemmD g) {
  __shared__ double As[16][64 + 1], Bs[16][64 + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += 16) {
    for (int idx = tid; idx < 64 * 16; idx += 256) {
      // walk the unit-stride dimension with consecutive threads
      int i, k;
      if (g.a_i == 1) { i = idx & 63; k = idx >> 6; } else { k = idx & 15; i = idx >> 4; }
      As[k][i] = (i0 + i < g.M && k0 + k < g.K) ? g.A[(long long)(i0 + i) * g.a_i + (long long)(k0 + k) * g.a_k] : 0.0;
      int j, kb;
      if (g.b_j == 1) { j = idx & 63; kb = idx >> 6; } else { kb = idx & 15; j = idx >> 4; }
      Bs[kb][j] = (j0 + j < g.N && k0 + kb < g.K) ? g.B[(long long)(k0 + kb) * g.b_k + (long long)(j0 + j) * g.b_j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { a[r] = As[k][ty * 4 + r]; b[r] = Bs[k][tx * 4 + r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
      if (i < g.M && j < g.N) {
        double *p = g.C + (long long)i * g.c_i + (long long)j * g.c_j;
        *p = (g.beta == 0.0 ? 0.0 : g.beta * *p) + g.alpha * acc[r][c];
      }
    }
}

constexpr int PB_NB = 32;   // panel width

// Cholesky of the nb x nb diagonal block at (k0, k0) of the row-major n x n matrix G (lower triangle), in place.
// info[0] is set to k0 + j + 1 when pivot j falls below 1e-12 of the column's squared norm diag0[k0 + j] (F not of
// full column rank to working precision: cond(F) beyond ~1e6).
__global__ void __launch_bounds__(1024) chol_diag_kernel(double *G, int n, int k0, int nb, const double *diag0, int *info) {
  __shared__ double S[PB_NB][PB_NB + 1];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  S[r][c] = (r < nb && c < nb) ? G[(long long)(k0 + r) * n + k0 + c] : 0.0;
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (r == j && c == j) {
      if (!(S[j][j] > 1e-12 * diag0[k0 + j])) { if (info[0] == 0) info[0] = k0 + j + 1; S[j][j] = 1.0; }
      S[j][j] = sqrt(S[j][j]);
    }
    __syncthreads();
    if (c == j && r > j) S[r][j] /= S[j][j];
    __syncthreads();
    if (r > j && c > j && c <= r) S[r][c] -= S[r][j] * S[c][j];
    __syncthreads();
  }
  if (r < nb && c < nb) G[(long long)(k0 + r) * n + k0 + c] = c <= r ? S[r][c] : 0.0;
}

// Panel below the diagonal block: row i of G[k1:, k0:k1] <- row i * L11^{-T} (one thread per row).
__global__ void __launch_bounds__(128) chol_panel_kernel(double *G, int n, int k0, int nb) {
  __shared__ double L[PB_NB][PB_NB + 1];
  for (int idx = threadIdx.x; idx < PB_NB * PB_NB; idx += blockDim.x) {
    const int r = idx >> 5, c = idx & 31;
    L[r][c] = (r < nb && c < nb) ? G[(long long)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int i = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[PB_NB];
  double *row = G + (long long)i * n + k0;
#pragma unroll
  for (int c = 0; c < PB_NB; ++c) x[c] = c < nb ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < PB_NB; ++c) {
    double s = x[c];
#pragma unroll
    for (int t = 0; t < PB_NB; ++t)
      if (t < c) s -= x[t] * L[c][t];
    x[c] = s / L[c][c];
  }
#pragma unroll
  for (int c = 0; c < PB_NB; ++c)
    if (c < nb) row[c] = x[c];
}

// Rows k0..k0+nb of W (n x D, row-major) <- L_kk^{-1} rows (one thread per column, forward substitution).
__global__ void __launch_bounds__(128) trsm_diag_kernel(const double *G, int n, double *W, int D, int k0, int nb) {
  __shared__ double L[PB_NB][PB_NB + 1];
  for (int idx = threadIdx.x; idx < PB_NB * PB_NB; idx += blockDim.x) {
    const int r = idx >> 5, c = idx & 31;
    L[r][c] = (r < nb && c < nb) ? G[(long long)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  double x[PB_NB];
#pragma unroll
  for (int c = 0; c < PB_NB; ++c) {
    double s = c < nb ? W[(long long)(k0 + c) * D + j] : 0.0;
#pragma unroll
    for (int t = 0; t < PB_NB; ++t)
      if (t < c) s -= L[c][t] * x[t];
    x[c] = s / L[c][c];
  }
#pragma unroll
  for (int c = 0; c < PB_NB; ++c)
    if (c < nb) W[(long long)(k0 + c) * D + j] = x[c];
}

__global__ void transpose_d_kernel(const double *F, double *W, int D, int r) {      // W[i][j] = F[j][i]
  __shared__ double t[32][33];
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y)
    if (j0 + y < D && i0 + threadIdx.x < r) t[y][threadIdx.x] = F[(long long)(j0 + y) * r + i0 + threadIdx.x];
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y)
    if (i0 + y < r && j0 + threadIdx.x < D) W[(long long)(i0 + y) * D + j0 + threadIdx.x] = t[threadIdx.x][y];
}

__global__ void copy_diag_kernel(const double *G, double *diag, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = G[(long long)i * n + i];
}

__global__ void cast_d2f_kernel(const double *in, float *out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

inline void launch_dgemm(const GemmD &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  dgemm_kernel<<<dim3((g.N + 63) / 64, (g.M + 63) / 64), 256, 0, st>>>(g);
}

// Host driver.  F: rows x cols fp64 row-major (host), P: rows x rows fp32 (host).  Returns cudaSuccess or the
// first CUDA error; *not_full_rank is the 1-based index of the first non-positive pivot, 0 if none.
inline cudaError_t build_projection_matrix(const double *F_host, int D, int r, float *P_host, int *not_full_rank) {
  double *dF = nullptr, *dG = nullptr, *dW = nullptr, *dP = nullptr, *dDiag = nullptr;
  float *dPf = nullptr;
  int *dInfo = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t e = cudaSuccess;
  auto done = [&](cudaError_t err) {
    cudaFree(dF); cudaFree(dG); cudaFree(dW); cudaFree(dP); cudaFree(dPf); cudaFree(dInfo); cudaFree(dDiag);
    if (st) cudaStreamDestroy(st);
    return err;
  };
#define PB_CK(call) do { e = (call); if (e != cudaSuccess) return done(e); } while (0)
  PB_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  PB_CK(cudaMalloc(&dF, sizeof(double) * D * r));
  PB_CK(cudaMalloc(&dG, sizeof(double) * r * r));
  PB_CK(cudaMalloc(&dW, sizeof(double) * r * D));
  PB_CK(cudaMalloc(&dP, sizeof(double) * D * D));
  PB_CK(cudaMalloc(&dPf, sizeof(float) * D * D));
  PB_CK(cudaMalloc(&dInfo, sizeof(int)));
  PB_CK(cudaMalloc(&dDiag, sizeof(double) * r));
  PB_CK(cudaMemsetAsync(dInfo, 0, sizeof(int), st));
  PB_CK(cudaMemcpyAsync(dF, F_host, sizeof(double) * D * r, cudaMemcpyHostToDevice, st));
  // 1. G = F^T F
  launch_dgemm(GemmD{dF, dF, dG, r, r, D, 1, r, r, 1, r, 1, 1.0, 0.0}, st);
  copy_diag_kernel<<<(r + 255) / 256, 256, 0, st>>>(dG, dDiag, r);
  // 2. Cholesky, lower triangle in place
  for (int k0 = 0; k0 < r; k0 += PB_NB) {
    const int nb = (r - k0 < PB_NB) ? r - k0 : PB_NB, k1 = k0 + nb;
    chol_diag_kernel<<<1, 1024, 0, st>>>(dG, r, k0, nb, dDiag, dInfo);
    if (k1 < r) {
      chol_panel_kernel<<<(r - k1 + 127) / 128, 128, 0, st>>>(dG, r, k0, nb);
      const double *L21 = dG + (long long)k1 * r + k0;
      launch_dgemm(GemmD{L21, L21, dG + (long long)k1 * r + k1, r - k1, r - k1, nb, r, 1, 1, r, r, 1, -1.0, 1.0}, st);
    }
  }
  // 3. W = L^{-1} F^T
  transpose_d_kernel<<<dim3((D + 31) / 32, (r + 31) / 32), dim3(32, 8), 0, st>>>(dF, dW, D, r);
  for (int k0 = 0; k0 < r; k0 += PB_NB) {
    const int nb = (r - k0 < PB_NB) ? r - k0 : PB_NB, k1 = k0 + nb;
    trsm_diag_kernel<<<(D + 127) / 128, 128, 0, st>>>(dG, r, dW, D, k0, nb);
    if (k1 < r)
      launch_dgemm(GemmD{dG + (long long)k1 * r + k0, dW + (long long)k0 * D, dW + (long long)k1 * D, r - k1, D, nb, r, 1, D, 1,
                         D, 1, -1.0, 1.0}, st);
  }
  // 4. P = W^T W
  launch_dgemm(GemmD{dW, dW, dP, D, D, r, 1, D, D, 1, D, 1, 1.0, 0.0}, st);
  cast_d2f_kernel<<<(unsigned)(((size_t)D * D + 255) / 256), 256, 0, st>>>(dP, dPf, (size_t)D * D);
  PB_CK(cudaGetLastError());
  PB_CK(cudaMemcpyAsync(P_host, dPf, sizeof(float) * D * D, cudaMemcpyDeviceToHost, st));
  PB_CK(cudaMemcpyAsync(not_full_rank, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  PB_CK(cudaStreamSynchronize(st));
#undef PB_CK
  return done(cudaSuccess);
}

// ---- least-squares dynamics fit on the device (fit_linear_dynamics, m_diffuser/dynamics/data_driven.py:107-121) ----
// Theta = argmin || [X U] Theta - X+ ||_F  (numpy lstsq in the reference), A = Theta[:n]^T, B = Theta[n:]^T.
// Normal equations in fp64: G = Z^T Z (k x k, k = n + m <= a few hundred), R = Z^T X+, G = L L^T,
// Theta = L^{-T} L^{-1} R computed as (L^{-1})^T (L^{-1} R) with the forward-substitution kernel above applied to R
// and to the identity.  cond(Z)^2 must stay below ~1e12 (the pivots are checked like the projector's); transition
// data with that conditioning would not identify (A, B) in the reference either.
__global__ void set_identity_kernel(double *M, int k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < k * k) M[idx] = (idx / k == idx % k) ? 1.0 : 0.0;
}

// Split-K Gram products for a tall-skinny Z: a block reduces 512 of the N rows, 32 at a time through shared memory,
// every thread keeping a strided set of the k x (k + n) outputs in registers, then adds its partial sums to the fp64
// results with atomics.  G[i][j] += sum_r Z[r][i] Z[r][j],  R[i][j] += sum_r Z[r][i] X+[r][j].
constexpr int GT_ROWS = 512, GT_CHUNK = 32, GT_OUT = 16;

__global__ void __launch_bounds__(256) gram_tall_kernel(const double *Z, int k, const double *Xn, int n, long long rows,
                                                         double *G, double *R) {
  extern __shared__ double gt_sm[];             // [GT_CHUNK rows][k + n]
  const int cols = k + n, n_out = k * cols;
  const long long r0 = (long long)blockIdx.x * GT_ROWS, r1 = (r0 + GT_ROWS < rows) ? r0 + GT_ROWS : rows;
  for (int pass0 = 0; pass0 < n_out; pass0 += 256 * GT_OUT) {
    double acc[GT_OUT];
    int oi[GT_OUT], oj[GT_OUT];
#pragma unroll
    for (int o = 0; o < GT_OUT; ++o) {
      acc[o] = 0.0;
      const int out = pass0 + o * 256 + (int)threadIdx.x;
      oi[o] = out < n_out ? out / cols : -1;
      oj[o] = out < n_out ? out - oi[o] * cols : 0;
    }
    for (long long rr = r0; rr < r1; rr += GT_CHUNK) {
      const int nr = (int)((r1 - rr < GT_CHUNK) ? r1 - rr : GT_CHUNK);
      __syncthreads();
      for (int idx = threadIdx.x; idx < nr * cols; idx += blockDim.x) {
        const int r = idx / cols, c = idx - r * cols;
        gt_sm[idx] = c < k ? Z[(rr + r) * k + c] : Xn[(rr + r) * n + (c - k)];
      }
      __syncthreads();
#pragma unroll
      for (int o = 0; o < GT_OUT; ++o) {
        if (oi[o] >= 0) {
          double a = acc[o];
          for (int r = 0; r < nr; ++r) a = fma(gt_sm[r * cols + oi[o]], gt_sm[r * cols + oj[o]], a);
          acc[o] = a;
        }
      }
    }
#pragma unroll
    for (int o = 0; o < GT_OUT; ++o) {
      if (oi[o] >= 0) {
        if (oj[o] < k) atomicAdd(&G[(long long)oi[o] * k + oj[o]], acc[o]);
        else atomicAdd(&R[(long long)oi[o] * n + (oj[o] - k)], acc[o]);
      }
    }
  }
}

// X (N x n), U (N x m), Xn (N x n): fp64 row-major HOST arrays.  A (n x n), B (n x m): fp64 row-major HOST outputs.
inline cudaError_t fit_linear_dynamics_device(const double *X, const double *U, const double *Xn, long long N, int n, int m,
                                              double *A_host, double *B_host, int *not_full_rank) {
  const int k = n + m;
  double *dZ = nullptr, *dXn = nullptr, *dG = nullptr, *dR = nullptr, *dLinv = nullptr, *dTheta = nullptr, *dDiag = nullptr;
  int *dInfo = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t e = cudaSuccess;
  auto done = [&](cudaError_t err) {
    cudaFree(dZ); cudaFree(dXn); cudaFree(dG); cudaFree(dR); cudaFree(dLinv); cudaFree(dTheta); cudaFree(dDiag); cudaFree(dInfo);
    if (st) cudaStreamDestroy(st);
    return err;
  };
#define FD_CK(call) do { e = (call); if (e != cudaSuccess) return done(e); } while (0)
  FD_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  FD_CK(cudaMalloc(&dZ, sizeof(double) * N * k));
  FD_CK(cudaMalloc(&dXn, sizeof(double) * N * n));
  FD_CK(cudaMalloc(&dG, sizeof(double) * k * k));
  FD_CK(cudaMalloc(&dR, sizeof(double) * k * n));
  FD_CK(cudaMalloc(&dLinv, sizeof(double) * k * k));
  FD_CK(cudaMalloc(&dTheta, sizeof(double) * k * n));
  FD_CK(cudaMalloc(&dDiag, sizeof(double) * k));
  FD_CK(cudaMalloc(&dInfo, sizeof(int)));
  FD_CK(cudaMemsetAsync(dInfo, 0, sizeof(int), st));
  FD_CK(cudaMemsetAsync(dG, 0, sizeof(double) * k * k, st));
  FD_CK(cudaMemsetAsync(dR, 0, sizeof(double) * k * n, st));
  // Z = [X | U]: two strided copies into the (N x k) device matrix
  FD_CK(cudaMemcpy2DAsync(dZ, sizeof(double) * k, X, sizeof(double) * n, sizeof(double) * n, (size_t)N, cudaMemcpyHostToDevice, st));
  FD_CK(cudaMemcpy2DAsync(dZ + n, sizeof(double) * k, U, sizeof(double) * m, sizeof(double) * m, (size_t)N, cudaMemcpyHostToDevice, st));
  FD_CK(cudaMemcpyAsync(dXn, Xn, sizeof(double) * N * n, cudaMemcpyHostToDevice, st));
  // G = Z^T Z, R = Z^T X+ (split over the rows)
  const size_t smem = sizeof(double) * GT_CHUNK * (size_t)(k + n);
  if (smem > 48 * 1024) FD_CK(cudaFuncSetAttribute(gram_tall_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_tall_kernel<<<(unsigned)((N + GT_ROWS - 1) / GT_ROWS), 256, smem, st>>>(dZ, k, dXn, n, N, dG, dR);
  copy_diag_kernel<<<(k + 255) / 256, 256, 0, st>>>(dG, dDiag, k);
  // G = L L^T
  for (int k0 = 0; k0 < k; k0 += PB_NB) {
    const int nb = (k - k0 < PB_NB) ? k - k0 : PB_NB, k1 = k0 + nb;
    chol_diag_kernel<<<1, 1024, 0, st>>>(dG, k, k0, nb, dDiag, dInfo);
    if (k1 < k) {
      chol_panel_kernel<<<(k - k1 + 127) / 128, 128, 0, st>>>(dG, k, k0, nb);
      const double *L21 = dG + (long long)k1 * k + k0;
      launch_dgemm(GemmD{L21, L21, dG + (long long)k1 * k + k1, k - k1, k - k1, nb, k, 1, 1, k, k, 1, -1.0, 1.0}, st);
    }
  }
  // Y = L^{-1} R (in place in dR) and Linv = L^{-1} I
  set_identity_kernel<<<(k * k + 255) / 256, 256, 0, st>>>(dLinv, k);
  for (int k0 = 0; k0 < k; k0 += PB_NB) {
    const int nb = (k - k0 < PB_NB) ? k - k0 : PB_NB, k1 = k0 + nb;
    trsm_diag_kernel<<<(n + 127) / 128, 128, 0, st>>>(dG, k, dR, n, k0, nb);
    trsm_diag_kernel<<<(k + 127) / 128, 128, 0, st>>>(dG, k, dLinv, k, k0, nb);
    if (k1 < k) {
      launch_dgemm(GemmD{dG + (long long)k1 * k + k0, dR + (long long)k0 * n, dR + (long long)k1 * n, k - k1, n, nb, k, 1, n, 1, n, 1, -1.0, 1.0}, st);
      launch_dgemm(GemmD{dG + (long long)k1 * k + k0, dLinv + (long long)k0 * k, dLinv + (long long)k1 * k, k - k1, k, nb, k, 1, k, 1, k, 1, -1.0, 1.0}, st);
    }
  }
  // Theta = Linv^T Y   (k x n)
  launch_dgemm(GemmD{dLinv, dR, dTheta, k, n, k, 1, k, n, 1, n, 1, 1.0, 0.0}, st);
  FD_CK(cudaGetLastError());
  std::vector<double> theta((size_t)k * n);
  FD_CK(cudaMemcpyAsync(theta.data(), dTheta, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st));
  FD_CK(cudaMemcpyAsync(not_full_rank, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  FD_CK(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) A_host[(size_t)i * n + j] = theta[(size_t)j * n + i];          // A = Theta[:n]^T
    for (int j = 0; j < m; ++j) B_host[(size_t)i * m + j] = theta[(size_t)(n + j) * n + i];    // B = Theta[n:]^T
  }
#undef FD_CK
  return done(cudaSuccess);
}

// ---- dynamics-violation metric on the device (ProjectionLoss.compute, m_diffuser/losses/__init__.py:161-186) ----
//   tau = [unnormalised states (x_0..x_{H-1}, x_{H-1} again) | unnormalised actions]   (B, Dc), Dc = (H+1) n + H m
//   residual = mean((tau - tau P)^2)
// One launch: a block owns RS_S samples x RS_N columns of tau P, builds its tau rows from the normalised
// trajectories on the fly, accumulates in fp32 like the reference's torch matmul, and adds its sum of squares to a
// fp64 scalar.
constexpr int RS_S = 16, RS_N = 128, RS_K = 32;

struct ResidualParams {
  const float *x;           // (B, H, T) normalised trajectories
  const float *P;           // (Dc, Dc) row-major
  const float *stats;       // obs_mean[n], obs_std[n], act_mean[m], act_std[m]
  double *out;              // sum of squared residuals (divide by B * Dc on the host)
  int B, H, n, m, T, Dc;
};

__device__ __forceinline__ float tau_elem(const ResidualParams &p, int b, int d) {
  const int ns = (p.H + 1) * p.n;
  if (d < ns) {
    int h = d / p.n;
    const int j = d - h * p.n;
    if (h == p.H) h = p.H - 1;                                   // duplicated last state (losses/__init__.py:153)
    return p.x[((size_t)b * p.H + h) * p.T + j] * p.stats[p.n + j] + p.stats[j];
  }
  const int e = d - ns, h = e / p.m, j = e - h * p.m;
  return p.x[((size_t)b * p.H + h) * p.T + p.n + j] * p.stats[2 * p.n + p.m + j] + p.stats[2 * p.n + j];
}

__global__ void __launch_bounds__(RS_N) residual_kernel(const ResidualParams p) {
  __shared__ float ts[RS_S][RS_K + 1];
  __shared__ float red[32];
  const int b0 = blockIdx.y * RS_S, c = blockIdx.x * RS_N + threadIdx.x;
  float acc[RS_S];
#pragma unroll
  for (int s = 0; s < RS_S; ++s) acc[s] = 0.f;
  for (int k0 = 0; k0 < p.Dc; k0 += RS_K) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < RS_S * RS_K; idx += RS_N) {
      const int s = idx / RS_K, k = idx - s * RS_K;
      ts[s][k] = (b0 + s < p.B && k0 + k < p.Dc) ? tau_elem(p, b0 + s, k0 + k) : 0.f;
    }
    __syncthreads();
    if (c < p.Dc) {
      const int kn = (p.Dc - k0 < RS_K) ? p.Dc - k0 : RS_K;
      for (int k = 0; k < kn; ++k) {
        const float w = __ldg(p.P + (size_t)(k0 + k) * p.Dc + c);
#pragma unroll
        for (int s = 0; s < RS_S; ++s) acc[s] = fmaf(ts[s][k], w, acc[s]);
      }
    }
  }
  float sq = 0.f;
  if (c < p.Dc) {
#pragma unroll
    for (int s = 0; s < RS_S; ++s)
      if (b0 + s < p.B) {
        const float d = tau_elem(p, b0 + s, c) - acc[s];
        sq = fmaf(d, d, sq);
      }
  }
  // block sum -> one fp64 atomic
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < RS_N / 32; ++w) t += (double)red[w];
    atomicAdd(p.out, t);
  }
}

}  // namespace dad
