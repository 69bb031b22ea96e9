// Measurement aid for bench.py's roofline: the ceiling of a random 256-byte-row gather.
//
// The gather kernels of the aggregation read rows of D = 64 floats at data-dependent positions from tables
// that mostly live in the 126 MB L2 (2.7 - 27 MB; one source of 179 MB does not), so the HBM copy peak is
// the wrong denominator for them.  This kernel issues the same access shape — 16 lanes x float4 per row, 8
// independent rows in flight per lane group, the launch geometry of gather_rows_fast_kernel — with NO index
// load in front of the row load (row ids come from an integer hash), no weights and no per-segment epilogue:
// what it reaches on a table of a given size is the rate the memory system can deliver such rows at.
#include "common.cuh"

namespace sg {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256, 1) row_gather_probe_kernel(float4 *__restrict__ out, const float4 *__restrict__ table,
                                                                  uint32_t n_rows, int reads_per_group, uint32_t seed) {
  const int lane = threadIdx.x & 15;
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const float4 *__restrict__ src = table + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t state = mix32(group * 0x9E3779B9u + seed);
  for (int i = 0; i < reads_per_group; i += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t row = __umulhi(mix32(state + (uint32_t)(i + u)), n_rows);
      v[u] = __ldg(src + (size_t)row * 16);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  out[(size_t)group * 16 + lane] = acc;
}

}  // namespace sg

using namespace sg;

extern "C" int sg_row_gather_probe(float *out, const float *table, int n_rows, int reads_per_group, int blocks,
                                   unsigned seed, sg_stream_t stream) {
  SG_REQUIRE(out && table && n_rows > 0 && reads_per_group > 0 && (reads_per_group & 7) == 0 && blocks > 0,
             "sg_row_gather_probe: bad arguments (reads_per_group must be a positive multiple of 8)");
  row_gather_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(out),
                                                                    reinterpret_cast<const float4 *>(table),
                                                                    (uint32_t)n_rows, reads_per_group, seed);
  SG_LAUNCHED("row_gather_probe_kernel");
  return SG_OK;
}
