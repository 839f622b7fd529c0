// scatter_prims.cu -- which primitive should carry the element -> CSR scatter on B200?
//
// The fused tangent kernel issues 224 RED.F64 sectors per element and its scatter phase alone needs 11.2 ms at 192^3
// (profiles/r01u_ko_sweep.txt): 141 G sectors/s.  Is that an SM-side (LSU) or an L2-side limit, and what do the
// alternatives reach on the SAME address pattern (hex8 lattice, node rows of 3 x 81 doubles, 4 runs of 6 doubles per
// (row, element))?
//   red       RED.E.ADD.F64 from registers, lane = column (24 of 32), warp walks the 24 rows of 5 elements  (= k_mat2 S2)
//   st        the same addresses with plain ST.E.64 (no read-modify-write at L2)
//   bulkred   cp.reduce.async.bulk.global.shared::cta.add.f64 per run (48 B, or 64 B zero-padded when misaligned): the
//             scatter leaves the LSU / L1 data pipe altogether; 32 lanes issue 3 runs each per element
//   bulkst    cp.async.bulk.global.shared::cta per run (plain bulk store, same sizes)
//   rowst     owner-computes output pattern: every node's 243 doubles written exactly once, coalesced ST.64 by a warp
//   rowbulk   the same with one 1936-byte bulk store per node (+ one scalar store for alignment)
//   red3      RED with x-merged runs (3 nodes = 9 doubles per run, the in-CTA combine of x-adjacent elements)
// Every mode is run with several CTAs-per-SM settings: if sectors/s scale with the number of SMs in use the limit is on
// the SM side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_prims scatter_prims.cu && ./scatter_prims [n=128]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t compact3(uint32_t x) {  // every third bit -> contiguous
  x &= 0x09249249u;
  x = (x ^ (x >> 2)) & 0x030c30c3u;
  x = (x ^ (x >> 4)) & 0x0300f00fu;
  x = (x ^ (x >> 8)) & 0xff0000ffu;
  x = (x ^ (x >> 16)) & 0x000003ffu;
  return x;
}
// hex8 local node -> lattice offset (Exodus ordering)
__constant__ int c_dx[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__constant__ int c_dy[8] = {0, 0, 1, 1, 0, 0, 1, 1};
__constant__ int c_dz[8] = {0, 0, 0, 0, 1, 1, 1, 1};

struct Geo { int n; int N; };  // n elements per axis, N = n + 1 nodes per axis (node rows padded to 27 neighbours)

__device__ __forceinline__ void elem_xyz(uint32_t e, int& ex, int& ey, int& ez) {
  ex = compact3(e); ey = compact3(e >> 1); ez = compact3(e >> 2);
}
__device__ __forceinline__ int64_t node_id(const Geo& g, int x, int y, int z) { return x + (int64_t)g.N * (y + (int64_t)g.N * z); }
// slot of entry (row dof d of node b, column dof dc of node k) for element at (ex, ey, ez)
__device__ __forceinline__ int64_t slot(const Geo& g, int ex, int ey, int ez, int b, int d, int k, int dc) {
  const int64_t nb = node_id(g, ex + c_dx[b], ey + c_dy[b], ez + c_dz[b]);
  const int ox = c_dx[k] - c_dx[b] + 1, oy = c_dy[k] - c_dy[b] + 1, oz = c_dz[k] - c_dz[b] + 1;
  return nb * 243 + d * 81 + ((oz * 3 + oy) * 3 + ox) * 3 + dc;
}

template <int MODE>  // 0 red, 1 st
__global__ void k_red(double* nz, Geo g, uint32_t ne) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int k = lane / 3, dc = lane % 3;
  for (uint32_t e0 = wid * 5; e0 < ne; e0 += nw * 5) {
    for (int el = 0; el < 5 && e0 + el < ne; ++el) {
      int ex, ey, ez;
      elem_xyz(e0 + el, ex, ey, ez);
      if (lane < 24) {
        int64_t a[24];
#pragma unroll
        for (int row = 0; row < 24; ++row) a[row] = slot(g, ex, ey, ez, row / 3, row % 3, k, dc);
#pragma unroll
        for (int row = 0; row < 24; ++row) {
          if (MODE == 0) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(nz + a[row]), "d"(1.0) : "memory");
          else asm volatile("st.global.f64 [%0], %1;" ::"l"(nz + a[row]), "d"(1.0) : "memory");
        }
      }
    }
  }
}

// x-merged runs: two x-adjacent elements scatter together; a row of the shared face sees 3-node runs (9 doubles)
__global__ void k_red3(double* nz, Geo g, uint32_t ne) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  // element PAIR (2 x 1 x 1): 12 nodes, 36 rows; row node at x-position 0,1,2 has 2,3,2 x-neighbours in the pair
  for (uint32_t p0 = wid * 2; p0 * 2 < ne; p0 += nw * 2) {
    for (int pl = 0; pl < 2 && (p0 + pl) * 2 < ne; ++pl) {
      int ex, ey, ez;
      elem_xyz((p0 + pl) * 2, ex, ey, ez);  // Morton: the x-neighbour is e + 1
      // lanes 0..26: column = (x-offset in {-1,0,1}) x dof, for each of the 4 (dy,dz) column lines x 36 rows
#pragma unroll 1
      for (int rn = 0; rn < 12; ++rn) {
        const int rx = rn % 3, ry = (rn / 3) % 2, rz = rn / 6;
        const int64_t nb = node_id(g, ex + rx, ey + ry, ez + rz);
#pragma unroll
        for (int line = 0; line < 4; ++line) {
          const int cy = line & 1, cz = line >> 1;
          const int oy = cy - ry + 1, oz = cz - rz + 1;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int ox = lane / 3, dc = lane % 3;  // ox in 0..2 -> x-offset -1..1
            const int cx = rx + ox - 1;
            if (lane < 9 && cx >= 0 && cx <= 2)
              asm volatile("red.global.add.f64 [%0], %1;" ::"l"(nz + nb * 243 + d * 81 + ((oz * 3 + oy) * 3 + ox) * 3 + dc), "d"(1.0) : "memory");
          }
        }
      }
    }
  }
}

template <int MODE>  // 0 bulk reduce, 1 bulk store
__global__ void k_bulk(double* nz, Geo g, uint32_t ne) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = sm + warp * 96 * 8;  // one 64-byte slot per run
  for (int i = lane; i < 96 * 8; i += 32) ws[i] = (i % 8 == 0 || i % 8 == 7) ? 0.0 : 1.0;  // [pad | 6 values | pad]
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t e = wid; e < ne; e += nw) {
    int ex, ey, ez;
    elem_xyz(e, ex, ey, ez);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int r = lane + 32 * j;           // run: row = r / 4, column line = r % 4
      const int row = r >> 2, line = r & 3;
      const int b = row / 3, d = row % 3;
      // the two x-adjacent column nodes of this line: local nodes with (dy, dz) = (line & 1, line >> 1)
      const int cy = line & 1, cz = line >> 1;
      const int k0 = (cz ? 4 : 0) + (cy ? 3 : 0);          // x = 0 node of the line: 0, 3, 4, 7
      int64_t s = slot(g, ex, ey, ez, b, d, k0, 0);
      const double* src = ws + r * 8 + 1;
      unsigned bytes = 48;
      if (s & 1) { s -= 1; src -= 1; bytes = 64; }           // 16-byte alignment: pad with the zero in front (and behind)
      const unsigned sa = (unsigned)__cvta_generic_to_shared(src);
      if ((s & 1) == 0 && (sa & 15) == 0) {
        if (MODE == 0)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(nz + s), "r"(sa), "r"(bytes) : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(nz + s), "r"(sa), "r"(bytes) : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>  // 0 warp-coalesced ST.64, 1 one bulk store per node
__global__ void k_rows(double* nz, int64_t nn) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = sm + warp * 244;
  for (int i = lane; i < 244; i += 32) ws[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = wid; n < nn; n += nw) {
    double* dst = nz + n * 243;
    if (MODE == 0) {
      for (int i = lane; i < 243; i += 32) dst[i] = 1.0;
    } else if (lane == 0) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(ws);
      if (n & 1) { dst[0] = 1.0; asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1936;" ::"l"(dst + 1), "r"(sa) : "memory"); }
      else { dst[242] = 1.0; asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1936;" ::"l"(dst), "r"(sa) : "memory"); }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float timeit(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 128;
  Geo g{n, n + 1};
  const uint32_t ne = (uint32_t)n * n * n;
  const int64_t nn = (int64_t)g.N * g.N * g.N;
  const size_t bytes = (size_t)(nn * 243 + 64) * 8;
  double* nz;
  CK(cudaMalloc(&nz, bytes));
  CK(cudaMemset(nz, 0, bytes));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("# %s, %d SMs, n = %d: %u elements, %lld node rows, target %.2f GB\n", prop.name, sms, n, ne, (long long)nn, bytes / 1e9);
  printf("%-10s %6s %8s %9s %12s %12s\n", "mode", "ctas/sm", "ms", "el/us", "Gsector/s", "GB/s(payload)");
  const double payload = 576.0 * 8;  // bytes of K_el entries per element
  auto report = [&](const char* name, int cps, float ms, double sectors_per_el, double payload_bytes, double units) {
    printf("%-10s %6d %8.3f %9.1f %12.1f %12.1f\n", name, cps, ms, units / ms / 1e3, sectors_per_el * units / ms / 1e6, payload_bytes * units / ms / 1e6);
  };
  CK(cudaFuncSetAttribute(k_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 96 * 64));
  CK(cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 96 * 64));
  for (int cps : {1, 2, 4, 8, 16}) {
    const int grid = sms * cps;
    report("red", cps, timeit([&] { k_red<0><<<grid, 64>>>(nz, g, ne); }), 224, payload, ne);
    report("st", cps, timeit([&] { k_red<1><<<grid, 64>>>(nz, g, ne); }), 224, payload, ne);
    report("red3", cps, timeit([&] { k_red3<<<grid, 64>>>(nz, g, ne); }), 0, 12 * 4 * 3 * 8 * 8.0 / 2, ne / 2.0);  // per pair: see kernel; sectors not modelled
    report("bulkred", cps, timeit([&] { k_bulk<0><<<grid, 64, 2 * 96 * 64>>>(nz, g, ne); }), 96 * 2.0, payload, ne);
    report("bulkst", cps, timeit([&] { k_bulk<1><<<grid, 64, 2 * 96 * 64>>>(nz, g, ne); }), 96 * 2.0, payload, ne);
    report("rowst", cps, timeit([&] { k_rows<0><<<grid, 64, 2 * 244 * 8>>>(nz, nn); }), 61, 1944, (double)nn);
    report("rowbulk", cps, timeit([&] { k_rows<1><<<grid, 64, 2 * 244 * 8>>>(nz, nn); }), 61, 1944, (double)nn);
  }
  // SM-side or L2-side limit?  One 512-thread CTA per SM (190 KB of dynamic shared memory keeps a second one out), on
  // all SMs and on half of them: an L2-side limit gives the same time, an SM-side limit doubles it.
  printf("# one 16-warp CTA per SM, all SMs vs half of them:\n");
  const int big = 190 * 1024;
  CK(cudaFuncSetAttribute(k_red<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_red<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  for (int div : {1, 2}) {
    report(div == 1 ? "red/all" : "red/half", 1, timeit([&] { k_red<0><<<sms / div, 512, big>>>(nz, g, ne); }), 224, payload, ne);
    report(div == 1 ? "st/all" : "st/half", 1, timeit([&] { k_red<1><<<sms / div, 512, big>>>(nz, g, ne); }), 224, payload, ne);
    report(div == 1 ? "bulkred/all" : "bulkred/half", 1, timeit([&] { k_bulk<0><<<sms / div, 512, big>>>(nz, g, ne); }), 192, payload, ne);
    report(div == 1 ? "rowst/all" : "rowst/half", 1, timeit([&] { k_rows<0><<<sms / div, 512, big>>>(nz, nn); }), 61, 1944, (double)nn);
  }
  CK(cudaFree(nz));
  return 0;
}
