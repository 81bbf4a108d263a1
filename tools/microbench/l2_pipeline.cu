// Dev micro-benchmark (not part of libslate_b200.so): does a two-pass strided transform keep its intermediate in
// L2 when the passes are pipelined slab by slab through a small, continuously reused scratch ring?
//   baseline : pass A (src -> tmp, full size) then pass B (tmp -> dst): 4 x bytes of HBM traffic
//   pipelined: ONE persistent kernel, tickets in a lag-1 interleaved order (A of slab k+1 between B of slab k),
//              intermediate in ring[slab % slots] (slots x slab bytes << L2): 2 x bytes of HBM traffic if L2 holds
// Access pattern = the cell FFT's (32 adjacent lines x 64 positions per CTA; pass A stride 16384 elements, pass B 256).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_pipeline l2_pipeline.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef double2 c128;
#ifndef ITEM_MLP
#define ITEM_MLP 8
#endif
#ifndef FENCE_MODE
#define FENCE_MODE 0
#endif

__device__ __forceinline__ c128 ld_hint(const c128* p, unsigned long long pol, int mode) {
  c128 v;
  if (mode == 0) {
    v = *p;
  } else if (mode == 1) {
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  } else {
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  }
  return v;
}
__device__ __forceinline__ void st_hint(c128* p, c128 v, unsigned long long pol, int mode) {
  if (mode < 2) *p = v;
  else asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol));
}

// one item: 32 lines x 64 positions; rows of `row_elems` elements; stage 0: stride sA, stage 1: stride 256
__device__ __forceinline__ void item(const c128* in, c128* out, int stage, int idx, long long row_elems, int mode_in,
                                     int mode_out, unsigned long long pol) {
  const int line = threadIdx.x & 31, j = threadIdx.x >> 5;
  long long base, stride;
  if (stage == 0) {
    stride = row_elems / 64;           // axis 0: 64 positions span the row
    base = (long long)idx * 32 + line;  // idx < stride / 32
  } else {
    stride = 256;                      // axis 1: 64 positions x 256 columns = 16384 elements per c0
    base = (long long)(idx / 8) * 16384 + (idx % 8) * 32 + line;
  }
  const c128* pi = in + base + (long long)j * stride;
  c128* po = out + base + (long long)j * stride;
  const long long s8 = 8 * stride;
#pragma unroll 1
  for (int h = 0; h < 8 / ITEM_MLP; ++h) {
    c128 x[ITEM_MLP];
#pragma unroll
    for (int t = 0; t < ITEM_MLP; ++t) x[t] = ld_hint(pi + t * s8, pol, mode_in);
#pragma unroll
    for (int t = 0; t < ITEM_MLP; ++t) {
      x[t].x += 1.0;
      st_hint(po + t * s8, x[t], pol, mode_out);
    }
    pi += ITEM_MLP * s8;
    po += ITEM_MLP * s8;
  }
}

__global__ void __launch_bounds__(256, 4) pass_kernel(const c128* in, c128* out, int stage, long long row_elems) {
  const int per_row = (int)(row_elems / 2048);
  const long long row = blockIdx.x / per_row;
  item(in + row * row_elems, out + row * row_elems, stage, blockIdx.x % per_row, row_elems, 0, 0, 0ull);
}

#ifndef PIPE_CTAS
#define PIPE_CTAS 8
#endif
__global__ void __launch_bounds__(256, PIPE_CTAS)
pipe_kernel(const c128* src, c128* ring, c128* dst, int slots, int rows, long long row_elems, unsigned* ticket,
            unsigned* done_a, unsigned* done_b, int hints) {
  __shared__ unsigned s_t;
  unsigned long long pol = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int per_row = (int)(row_elems / 2048);
  const long long total = 2LL * rows * per_row;
  int known_a = -1, known_b = -1;
  if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
  __syncthreads();
  for (;;) {
    const long long t = s_t;
    __syncthreads();
    if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);  // next ticket, in flight behind this item
    if (t >= total) break;
    // phase 0: per_row tickets (A_0); phases 1..rows-1: 2 per_row (B_{p-1}, A_p interleaved); phase rows: B_{rows-1}
    int stage, row, idx;
    if (t < per_row) {
      stage = 0; row = 0; idx = (int)t;
    } else {
      const long long u = t - per_row;
      const int p = (int)(u / (2 * per_row)) + 1;
      const int w = (int)(u % (2 * per_row));
      if (p < rows) {
        stage = (w & 1) ? 0 : 1;
        idx = w >> 1;
        row = stage ? p - 1 : p;
      } else {  // the tail phase holds only B items, two per pair slot
        stage = 1; row = rows - 1; idx = w;
        if (idx >= per_row) { __syncthreads(); continue; }
      }
    }
    c128* slot = ring + (long long)(row % slots) * row_elems;
    const unsigned target = (unsigned)per_row * (FENCE_MODE == 1 ? 8u : 1u);
    if (stage == 0 ? (row >= slots && row - slots > known_b) : (row > known_a)) {  // uniform over the CTA
      if (threadIdx.x == 0) {
        const unsigned* f = stage == 0 ? done_b + row - slots : done_a + row;
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        } while (v < target);
      }
      __syncthreads();
      if (stage == 0) known_b = row - slots; else known_a = row;
    }
    if (stage == 0) item(src + row * row_elems, slot, 0, idx, row_elems, hints ? 2 : 1, 0, pol);
    else item(slot, dst + row * row_elems, 1, idx, row_elems, 1, hints ? 2 : 0, pol);
    if (FENCE_MODE == 1) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(stage == 0 ? done_a + row : done_b + row) : "memory");
      __syncthreads();
    } else {
      __syncthreads();
      if (threadIdx.x == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(stage == 0 ? done_a + row : done_b + row) : "memory");
    }
  }
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (long long row_mb : {8LL, 16LL, 32LL}) {
    const long long row_elems = row_mb << 16;  // 16 B elements
    const size_t bytes = (size_t)rows * row_elems * 16;
    c128 *src, *tmp, *dst, *ring;
    unsigned* ctr;
    cudaMalloc(&src, bytes);
    cudaMalloc(&tmp, bytes);
    cudaMalloc(&dst, bytes);
    cudaMalloc(&ring, (size_t)8 * row_elems * 16);
    cudaMalloc(&ctr, (2 * rows + 64) * sizeof(unsigned));
    cudaMemset(src, 0, bytes);
    const int per_row = (int)(row_elems / 2048);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      pass_kernel<<<rows * per_row, 256>>>(src, tmp, 0, row_elems);
      pass_kernel<<<rows * per_row, 256>>>(tmp, dst, 1, row_elems);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("row %lld MB x %d rows: two passes %.3f ms  (%.0f GB/s of src read + dst write)\n", row_mb, rows, ms,
           2.0 * bytes / ms / 1e6);
    double sum = 0;
    // chunked launches: A on c rows -> ring, B ring -> dst (stream order = the dependency); kernels see L2-resident scratch?
    for (int c : {1, 2, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int r0 = 0; r0 < rows; r0 += c) {
          pass_kernel<<<c * per_row, 256>>>(src + (size_t)r0 * row_elems, ring, 0, row_elems);
          pass_kernel<<<c * per_row, 256>>>(ring, dst + (size_t)r0 * row_elems, 1, row_elems);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("   chunked launches, %d rows (%lld MB scratch): %.3f ms  (%.0f GB/s)\n", c, c * row_mb, ms, 2.0 * bytes / ms / 1e6);
    }
    for (int slots : {3, 4}) {
      for (int hints = 0; hints < 2; ++hints) {
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(ctr, 0, (2 * rows + 64) * sizeof(unsigned));
          cudaEventRecord(e0);
          pipe_kernel<<<148 * PIPE_CTAS, 256>>>(src, ring, dst, slots, rows, row_elems, ctr, ctr + 32, ctr + 32 + rows, hints);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
        }
        cudaError_t err = cudaGetLastError();
        printf("   pipelined slots %d (ring %lld MB) hints %d: %.3f ms  (%.0f GB/s)  %s\n", slots, slots * row_mb, hints,
               ms, 2.0 * bytes / ms / 1e6, cudaGetErrorString(err));
      }
    }
    // check: dst = src + 2 on a sample
    c128 h[4];
    cudaMemcpy(h, dst + (size_t)(rows - 1) * row_elems + 12345, sizeof(h), cudaMemcpyDeviceToHost);
    for (auto& v : h) sum += v.x;
    printf("   check (expect 8): %.1f\n", sum);
    cudaFree(src); cudaFree(tmp); cudaFree(dst); cudaFree(ring); cudaFree(ctr);
  }
  return 0;
}
