// Stand-alone check of the tensor-core contraction against the FP32-pipe contraction and a
// float64 reference, through the C ABI only (no torch).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -o build/tc_contract_check \
//        scripts/dev/tc_contract_check.cu -L bskit_b200 -lbskit_b200 -Xlinker -rpath='$ORIGIN/../bskit_b200'
//   ./build/tc_contract_check [log2_cells_small=20] [log2_cells_big=27] [nrows=40]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "bskit_b200.h"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(2); } } while (0)
#define BK(x) do { int r_ = (x); if (r_ != 0) { printf("bsk error %d (%s) at %s:%d\n", r_, bsk_last_error(), __FILE__, __LINE__); fflush(stdout); exit(3); } } while (0)

// smooth-ish pseudo-random shell fields: a few plane waves per row plus hashed noise
__global__ void fill_kernel(float* f, int64_t n, int row, float amp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + (uint64_t)(row + 1) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31; h *= 0x94D049BB133111EBull; h ^= h >> 29;
    const float u1 = ((h & 0xFFFFFF) + 0.5f) / 16777216.f, u2 = (((h >> 24) & 0xFFFFFF) + 0.5f) / 16777216.f;
    const float g = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    const float w = sinf(0.001f * (row + 3) * (float)(i % 100003)) + cosf(0.0007f * (row + 1) * (float)(i % 70001));
    f[i] = amp * (g + 0.7f * w);
  }
}
__global__ void ref_kernel(const float* const* rows, const int* tri, int ntri, int64_t n, double* out) {
  const int t = blockIdx.x;
  const float *a = rows[tri[3 * t]], *b = rows[tri[3 * t + 1]], *c = rows[tri[3 * t + 2]];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)a[i] * (double)b[i] * (double)c[i];
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[t] = sh[0];
}

static void run_mode(const char* mode, bsk_cplan* cp, const std::vector<void*>& ptrs, int64_t n, int ntri, double* d_out,
                     std::vector<double>& out, float* ms) {
  setenv("BSK_CONTRACT_MODE", mode, 1);
  const int32_t joff[3] = {0, 0, 0};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  BK(bsk_contract(cp, (const void* const*)ptrs.data(), BSK_F32, BSK_F32, n, 1, joff, d_out, nullptr));  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  BK(bsk_contract(cp, (const void* const*)ptrs.data(), BSK_F32, BSK_F32, n, 1, joff, d_out, nullptr));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(ms, e0, e1));
  out.resize(ntri);
  CK(cudaMemcpy(out.data(), d_out, sizeof(double) * ntri, cudaMemcpyDeviceToHost));
}

int main(int argc, char** argv) {
  const int lg_small = argc > 1 ? atoi(argv[1]) : 20, lg_big = argc > 2 ? atoi(argv[2]) : 27;
  const int R = argc > 3 ? atoi(argv[3]) : 40;
  std::vector<int32_t> tri;
  for (int a = 0; a < R; ++a)
    for (int b = a; b < R; ++b)
      for (int c = b; c < R; ++c)
        if (c <= a + b + 2) { tri.push_back(c); tri.push_back(b); tri.push_back(a); }   // k1 >= k2 >= k3 order
  const int ntri = (int)tri.size() / 3;
  printf("rows %d, triangles %d\n", R, ntri);
  bsk_cplan* cp = nullptr;
  BK(bsk_cplan_create(&cp, ntri, tri.data(), R, 1));
  for (int pass = 0; pass < 2; ++pass) {
    const int lg = pass ? lg_big : lg_small;
    if (lg <= 0) continue;
    const int64_t n = (int64_t)1 << lg;
    std::vector<void*> ptrs(R);
    for (int r = 0; r < R; ++r) {
      CK(cudaMalloc(&ptrs[r], sizeof(float) * n));
      fill_kernel<<<1024, 256>>>((float*)ptrs[r], n, r, 1.0f + 0.05f * r);
    }
    CK(cudaDeviceSynchronize());
    double* d_out; CK(cudaMalloc(&d_out, sizeof(double) * ntri));
    std::vector<double> packed, tcv, ref;
    float ms1 = 0, ms2 = 0;
    run_mode("1", cp, ptrs, n, ntri, d_out, packed, &ms1);
    run_mode("2", cp, ptrs, n, ntri, d_out, tcv, &ms2);
    printf("cells 2^%d: FP32-pipe %.3f ms, tensor-core %.3f ms\n", lg, ms1, ms2);
    if (!pass) {
      void** d_rows; int* d_tri;
      CK(cudaMalloc(&d_rows, sizeof(void*) * R)); CK(cudaMemcpy(d_rows, ptrs.data(), sizeof(void*) * R, cudaMemcpyHostToDevice));
      CK(cudaMalloc(&d_tri, sizeof(int) * 3 * ntri)); CK(cudaMemcpy(d_tri, tri.data(), sizeof(int) * 3 * ntri, cudaMemcpyHostToDevice));
      ref_kernel<<<ntri, 256>>>((const float* const*)d_rows, d_tri, ntri, n, d_out);
      CK(cudaDeviceSynchronize());
      ref.resize(ntri);
      CK(cudaMemcpy(ref.data(), d_out, sizeof(double) * ntri, cudaMemcpyDeviceToHost));
    } else {
      ref = packed;
    }
    double mx = 0, e1 = 0, e2 = 0, rel2 = 0; int worst = 0;
    for (int t = 0; t < ntri; ++t) mx = fmax(mx, fabs(ref[t]));
    for (int t = 0; t < ntri; ++t) {
      e1 = fmax(e1, fabs(packed[t] - ref[t]));
      if (fabs(tcv[t] - ref[t]) > e2) { e2 = fabs(tcv[t] - ref[t]); worst = t; }
      if (fabs(ref[t]) > 1e-3 * mx) rel2 = fmax(rel2, fabs(tcv[t] - ref[t]) / fabs(ref[t]));
    }
    printf("  vs %s: max|ref| %.4e;  FP32-pipe max err %.3e (%.2e of max);  tensor-core max err %.3e (%.2e of max), max rel (|ref|>1e-3 max) %.2e\n",
           pass ? "FP32-pipe result" : "float64 reference", mx, e1, e1 / mx, e2, e2 / mx, rel2);
    printf("  worst triangle %d (%d,%d,%d): ref %.9e tc %.9e packed %.9e\n", worst, tri[3 * worst], tri[3 * worst + 1], tri[3 * worst + 2],
           ref[worst], tcv[worst], packed[worst]);
    {
      double s1 = 0, s2 = 0, b2 = 0; int nself = 0; double sself = 0, bself = 0;
      for (int t = 0; t < ntri; ++t) {
        const double d = tcv[t] - ref[t];
        s2 += d * d; s1 += (packed[t] - ref[t]) * (packed[t] - ref[t]);
        b2 += d * (ref[t] > 0 ? 1 : -1);
        if (tri[3 * t + 1] == tri[3 * t + 2]) { ++nself; sself += d * d; bself += d * (ref[t] > 0 ? 1 : -1); }
      }
      printf("  rms err / max: FP32-pipe %.2e, tensor-core %.2e; mean signed (toward |ref| growth) %.2e of max; self-pair triangles (%d): rms %.2e mean %.2e\n",
             sqrt(s1 / ntri) / mx, sqrt(s2 / ntri) / mx, b2 / ntri / mx, nself, sqrt(sself / nself) / mx, bself / nself / mx);
    }
    for (int t : {0, 1, ntri / 2, ntri - 1}) printf("  t=%d ref %.9e tc %.9e\n", t, ref[t], tcv[t]);
    for (int r = 0; r < R; ++r) CK(cudaFree(ptrs[r]));
    CK(cudaFree(d_out));
    fflush(stdout);
  }
  bsk_cplan_destroy(cp);
  return 0;
}
