// Cloud-in-cell mass assignment of particles onto a periodic mesh (sm_100a).
//
// Replaces the step before the bispectrum path for particle inputs:
// `catalog.to_mesh(Nmesh, BoxSize, window='cic')` + paint in the reference's drivers
// (scripts/measure/measure_bs_fast.py:209-217, measure_subbox_bs_fast.py:225-233), which
// nbodykit/pmesh execute on the CPU.  Convention (pmesh's): mesh points sit at integer
// multiples of BoxSize/Nmesh; a particle at grid coordinate g = x * Nmesh / BoxSize gives weight
// prod_axis (1 - |g_axis - i_axis|) to the 8 mesh points around it, periodically wrapped.
//
// HBM/atomic bound: 12 bytes read per particle (float32 positions) + 8 float32 REDs that mostly
// hit L2.  Particles are processed in the order given; sorted inputs coalesce best.
#include "common.cuh"

namespace bsk {

template <typename T>
__global__ void __launch_bounds__(256)
paint_cic_kernel(const T* __restrict__ pos, int64_t npart, int n, double sx, double sy, double sz,
                 float* __restrict__ mesh) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npart; p += stride) {
    // grid coordinates in float64: the cell index must not depend on float32 rounding of x*N/L
    const double gx = (double)pos[3 * p + 0] * sx, gy = (double)pos[3 * p + 1] * sy, gz = (double)pos[3 * p + 2] * sz;
    const double fx = floor(gx), fy = floor(gy), fz = floor(gz);
    const float dx = (float)(gx - fx), dy = (float)(gy - fy), dz = (float)(gz - fz);
    int ix = (int)fmod(fx, (double)n), iy = (int)fmod(fy, (double)n), iz = (int)fmod(fz, (double)n);
    if (ix < 0) ix += n;
    if (iy < 0) iy += n;
    if (iz < 0) iz += n;
    const int jx = ix + 1 == n ? 0 : ix + 1, jy = iy + 1 == n ? 0 : iy + 1, jz = iz + 1 == n ? 0 : iz + 1;
    const float wx0 = 1.f - dx, wy0 = 1.f - dy, wz0 = 1.f - dz;
    const size_t nn = (size_t)n;
    float* r00 = mesh + ((size_t)ix * nn + iy) * nn;
    float* r01 = mesh + ((size_t)ix * nn + jy) * nn;
    float* r10 = mesh + ((size_t)jx * nn + iy) * nn;
    float* r11 = mesh + ((size_t)jx * nn + jy) * nn;
    atomicAdd(r00 + iz, wx0 * wy0 * wz0);
    atomicAdd(r00 + jz, wx0 * wy0 * dz);
    atomicAdd(r01 + iz, wx0 * dy * wz0);
    atomicAdd(r01 + jz, wx0 * dy * dz);
    atomicAdd(r10 + iz, dx * wy0 * wz0);
    atomicAdd(r10 + jz, dx * wy0 * dz);
    atomicAdd(r11 + iz, dx * dy * wz0);
    atomicAdd(r11 + jz, dx * dy * dz);
  }
}

}  // namespace bsk

extern "C" int bsk_paint_cic(const void* pos, int precision, int64_t npart, int nmesh,
                             const double boxsize[3], float* mesh, void* cuda_stream) {
  using namespace bsk;
  BSK_REQUIRE(pos && mesh && boxsize, "bsk_paint_cic: null argument");
  BSK_REQUIRE(npart >= 0 && nmesh >= 2 && nmesh <= 8192, "bsk_paint_cic: bad npart / nmesh");
  BSK_REQUIRE(precision == BSK_F32 || precision == BSK_F64, "bsk_paint_cic: bad precision");
  BSK_REQUIRE(boxsize[0] > 0 && boxsize[1] > 0 && boxsize[2] > 0, "bsk_paint_cic: BoxSize must be positive");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  BSK_CUDA(cudaMemsetAsync(mesh, 0, sizeof(float) * (size_t)nmesh * nmesh * nmesh, st));
  if (npart == 0) return BSK_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<int64_t>((npart + 255) / 256, (int64_t)sms * 16);
  const double sx = nmesh / boxsize[0], sy = nmesh / boxsize[1], sz = nmesh / boxsize[2];
  if (precision == BSK_F32)
    paint_cic_kernel<float><<<grid, 256, 0, st>>>((const float*)pos, npart, nmesh, sx, sy, sz, mesh);
  else
    paint_cic_kernel<double><<<grid, 256, 0, st>>>((const double*)pos, npart, nmesh, sx, sy, sz, mesh);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  return BSK_OK;
}
