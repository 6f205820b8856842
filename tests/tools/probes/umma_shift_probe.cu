// Probe (test infrastructure, not part of the library): can one 128B-swizzled K-major operand buffer in shared memory be
// read by tcgen05.mma at an arbitrary ROW offset?  That is what operand reuse across convolution taps needs (DESIGN.md
// section 7, item 1): the A tile of tap (dy,dx) is the same pixel rows shifted by dy*(W+2)+dx.
//
// One CTA.  A_full: R = 192 rows x 64 fp16 (128 B per row) written with the library's swizzle (row r, 16-byte chunk c at
// r*128 + ((c ^ r) & 7)*16, i.e. the XOR uses the ABSOLUTE row index).  B: 64 rows x 64 fp16, same layout.
// For every shift d the MMA D[128 x 64] = A_full[d .. d+128) * B^T is issued with the A descriptor's start address
// advanced by d*128 bytes, under three conventions for the descriptor's base-offset field (bits 49..51):
//   mode 0: base_offset = 0
//   mode 1: base_offset = (start_address >> 7) & 7        (what the PTX ISA text describes)
//   mode 2: base_offset = (8 - ((start_address >> 7) & 7)) & 7
// and compared with the exact integer result.  Prints max |error| per (shift, mode); a mode whose errors are all 0 is the
// addressing rule to use.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I coclr_b200/csrc -o /tmp/umma_shift_probe \
//        tests/tools/probes/umma_shift_probe.cu && /tmp/umma_shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

using namespace coclr;

static constexpr int kRows = 192, kK = 64, kN = 64, kM = 128;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __half* a_full, const __half* b, float* out, int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // kRows * 128 B
  uint8_t* sb = smem + kRows * 128;         // kN * 128 B (1024-aligned: 192*128 = 24576)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * 128);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kRows * 8; i += 128) {   // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sa + swz128_offset(r, c)) = *reinterpret_cast<const uint4*>(a_full + r * kK + c * 8);
  }
  for (int i = tid; i < kN * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sb + swz128_offset(r, c)) = *reinterpret_cast<const uint4*>(b + r * kK + c * 8);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<64>(holder);
  fence_proxy_async_smem();                  // generic-proxy smem writes -> async-proxy (tensor core) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(sa) + (uint32_t)shift * 128u;
    uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
    uint32_t bo = (a_addr >> 7) & 7u;
    if (mode == 0) bo = 0;
    if (mode == 2) bo = (8u - bo) & 7u;
    adesc |= (uint64_t)bo << 49;
    const uint64_t bdesc = make_smem_desc(smem_u32(sb), 16, 1024);
    const uint32_t idesc = make_idesc(0u, 0u, 0u, 0u, kM, kN);
    for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (tid & 31)) * kN + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  std::vector<__half> ha(kRows * kK), hb(kN * kK);
  std::vector<int> ia(kRows * kK), ib(kN * kK);
  srand(1);
  for (int i = 0; i < kRows * kK; ++i) { ia[i] = rand() % 7 - 3; ha[i] = __float2half((float)ia[i]); }
  for (int i = 0; i < kN * kK; ++i) { ib[i] = rand() % 5 - 2; hb[i] = __float2half((float)ib[i]); }
  __half *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, kM * kN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + kN * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 8, 16, 1, 2, 3, 5, 7, 9, 13, 34, 35, 63};
  std::vector<float> hout(kM * kN);
  for (int shift : shifts) {
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemset(dout, 0, kM * kN * 4);
      probe_kernel<<<1, 128, smem>>>(da, db, dout, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("shift %2d mode %d: CUDA error %s\n", shift, mode, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(hout.data(), dout, kM * kN * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          long ref = 0;
          for (int k = 0; k < kK; ++k) ref += (long)ia[(m + shift) * kK + k] * ib[n * kK + k];
          const double d = fabs((double)hout[m * kN + n] - (double)ref);
          if (d > worst) worst = d;
        }
      printf("shift %2d  base_offset mode %d  max|err| %g\n", shift, mode, worst);
    }
  }
  return 0;
}
