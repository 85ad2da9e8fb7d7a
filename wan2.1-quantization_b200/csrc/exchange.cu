// Head-group exchange of the sequence-parallel attention over NVLink peer memory.
//
// Replaces the all-to-alls of the reference's Ulysses attention (ViDiT-Q/examples/Wan2.1/wan/distributed/
// xdit_context_parallel.py:149-192 -> xfuser/yunchang SeqAllToAll4D: four c10d all_to_all_single calls per block, each
// with a permuting copy on either side).  Every rank holds its peers' receive buffers as plain device pointers (CUDA
// peer mappings of one symmetric allocation per rank); ONE launch reads the rank's q|k|v (or attention output) where the
// producing kernel left it and stores every destination's slice straight into that destination's buffer, in the layout
// the attention kernel (or the output projection's quantizer) reads.  No staging copies, no NCCL call; the ranks meet
// at one signal-pad barrier before the data is consumed.
#include "common.cuh"

namespace b200q {

constexpr int kMaxScatter = 48;          // 8 ranks x (q, k, v) x 2 CFG branches

struct ScatterArgs {
  const uint4* src[kMaxScatter];
  uint4* dst[kMaxScatter];
  long long src_pitch[kMaxScatter];      // in 16-byte units (q / k are contiguous, v is a column slice of the q|k|v GEMM output)
  long long dst_pitch;
  int rows, vec_per_row, n;
};

// blockIdx.y = message; 16 bytes per thread and step, consecutive threads on consecutive 16-byte words of a row, so a
// warp stores 512 contiguous bytes to the peer (full NVLink write packets).
__global__ void __launch_bounds__(256) scatter_rows_kernel(const __grid_constant__ ScatterArgs a) {
  const int e = blockIdx.y;
  const uint4* __restrict__ s = a.src[e];
  uint4* __restrict__ d = a.dst[e];
  const long long total = (long long)a.rows * a.vec_per_row, sp = a.src_pitch[e];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / a.vec_per_row;
    const int c = (int)(i - r * a.vec_per_row);
    d[r * a.dst_pitch + c] = ldg_stream16(s + r * sp + c);
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_scatter_rows(const void* const* src, void* const* dst, int n, int64_t rows, int64_t row_bytes,
                                  const int64_t* src_pitch_bytes, int64_t dst_pitch_bytes, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(src && dst && src_pitch_bytes && n >= 0 && n <= kMaxScatter, B200Q_ERR_BAD_ARG, "scatter_rows: 0..%d messages", kMaxScatter);
  B200Q_REQUIRE(rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0 &&
                    dst_pitch_bytes >= row_bytes && rows < (1ll << 31) && row_bytes < (1ll << 31),
                B200Q_ERR_BAD_ARG, "scatter_rows: rows of a multiple of 16 bytes, pitches multiples of 16 and >= the row");
  if (n == 0 || rows == 0) return B200Q_OK;
  ScatterArgs a{};
  for (int i = 0; i < n; ++i) {
    B200Q_REQUIRE(src[i] && dst[i] && aligned(src[i], 16) && aligned(dst[i], 16) && src_pitch_bytes[i] % 16 == 0 &&
                      src_pitch_bytes[i] >= row_bytes,
                  B200Q_ERR_BAD_ARG, "scatter_rows: message %d: null / not 16-byte aligned / bad source pitch", i);
    a.src[i] = (const uint4*)src[i];
    a.dst[i] = (uint4*)dst[i];
    a.src_pitch[i] = src_pitch_bytes[i] / 16;
  }
  a.dst_pitch = dst_pitch_bytes / 16;
  a.rows = (int)rows; a.vec_per_row = (int)(row_bytes / 16); a.n = n;
  const long long total = rows * (row_bytes / 16);
  long long bx = (total + 256 * 4 - 1) / (256 * 4);               // >= 4 words per thread
  const long long cap = (4LL * sm_count() + n - 1) / n;            // about four CTAs per SM over all messages
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  scatter_rows_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(a);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}
