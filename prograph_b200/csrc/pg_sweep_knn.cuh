// kNN-specialised variant of the fused Hamming sweep: TM own rows per consumer thread.
//
// Same ring, same producer, same list/threshold logic as sweep_kernel<.., MODE_KNN, ..>
// (pg_sweep.cuh); the difference is that every 128-bit shared-memory load of a stream row
// feeds TM own rows, halving (TM=2) the LDS traffic and doubling the independent LOP3
// chains per thread.  Chosen per shape from measurements (profiles/).
#pragma once
#include "pg_sweep.cuh"

namespace pg {

template <int P, int W, int TM>
__device__ __forceinline__ void ham_rows(const uint32_t (&q)[TM][P * W], const uint32_t* __restrict__ col,
                                         int (&d)[TM]) {
  static_assert(W % 4 == 0, "multi-row variant is built for W in {4, 8}");
  uint32_t m[TM][W];
  const uint4* c4 = reinterpret_cast<const uint4*>(col);
#pragma unroll
  for (int p = 0; p < P; ++p) {
#pragma unroll
    for (int h = 0; h < W / 4; ++h) {
      const uint4 v = c4[p * (W / 4) + h];
      const int w = h * 4;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        if (p == 0) {
          m[i][w + 0] = q[i][w + 0] ^ v.x;
          m[i][w + 1] = q[i][w + 1] ^ v.y;
          m[i][w + 2] = q[i][w + 2] ^ v.z;
          m[i][w + 3] = q[i][w + 3] ^ v.w;
        } else {
          m[i][w + 0] |= q[i][p * W + w + 0] ^ v.x;
          m[i][w + 1] |= q[i][p * W + w + 1] ^ v.y;
          m[i][w + 2] |= q[i][p * W + w + 2] ^ v.z;
          m[i][w + 3] |= q[i][p * W + w + 3] ^ v.w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) s += __popc(m[i][w]);
    d[i] = s;
  }
}

template <int P, int W, int TM, int MINB>
__global__ void __launch_bounds__(kSweepThreads, MINB) sweep_knn_kernel(const __grid_constant__ SweepParams prm) {
  constexpr int BN = TileCols<W>::value;
  constexpr int COLW = P * W;
  constexpr uint32_t STAGE_BYTES = BN * COLW * 4;
  constexpr int ROWS_CTA = kConsumers * TM;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* stage_mem = reinterpret_cast<uint32_t*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * STAGE_BYTES);
  uint64_t* empty = full + kStages;
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(empty + kStages + kMaxLutWords / 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int n_items = prm.n_rowblocks * prm.n_splits;   // n_rowblocks counts ROWS_CTA-row blocks here
  int stage = 0;
  uint32_t phase = 0;

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int split = item / prm.n_rowblocks;
        const int t0 = split * prm.tiles_per_split;
        const int t1 = min(t0 + prm.tiles_per_split, prm.n_tiles);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          bulk_g2s(stage_mem + stage * (BN * COLW), prm.str + static_cast<size_t>(t) * BN * COLW, STAGE_BYTES,
                   &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int split = item / prm.n_rowblocks;
    const int rb = item - split * prm.n_rowblocks;
    long long r[TM];
    bool valid[TM];
    uint32_t q[TM][COLW];
    unsigned tau[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      r[i] = static_cast<long long>(rb) * ROWS_CTA + i * kConsumers + tid;
      valid[i] = r[i] < prm.rows;
      const uint32_t* src = prm.own + static_cast<size_t>(prm.own_row0 + (valid[i] ? r[i] : 0)) * COLW;
#pragma unroll
      for (int j = 0; j < COLW; ++j) q[i][j] = valid[i] ? __ldg(src + j) : 0u;
      tau[i] = valid[i] ? 0xffffffffu : 0u;
      for (int j = 0; j < prm.k1; ++j) lists[(i * prm.k1 + j) * kConsumers + tid] = ~0ull;
    }

    const int t0 = split * prm.tiles_per_split;
    const int t1 = min(t0 + prm.tiles_per_split, prm.n_tiles);
    for (int t = t0; t < t1; ++t) {
      mbar_wait(&full[stage], phase);
      const uint32_t* tile = stage_mem + stage * (BN * COLW);
      const long long col0 = static_cast<long long>(t) * BN;
      const int ncols = static_cast<int>(min(static_cast<long long>(BN), prm.str_rows - col0));
      for (int c = 0; c < ncols; ++c) {
        int d[TM];
        ham_rows<P, W, TM>(q, tile + c * COLW, d);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          if (static_cast<unsigned>(d[i]) < tau[i]) {
            const unsigned long long key = (static_cast<unsigned long long>(static_cast<unsigned>(d[i])) << 32) |
                                           static_cast<unsigned>(col0 + c);
            tau[i] = knn_insert(lists + i * prm.k1 * kConsumers + tid, kConsumers, prm.k1, key);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
      if (valid[i]) {
        unsigned long long* dst = prm.part + static_cast<size_t>(split) * prm.k1 * prm.rows + r[i];
        for (int j = 0; j < prm.k1; ++j)
          dst[static_cast<size_t>(j) * prm.rows] = lists[(i * prm.k1 + j) * kConsumers + tid];
      }
    }
  }
}

template <int P, int W, int TM, int MINB>
int launch_knn_variant(const SweepParams& prm, size_t list_bytes, cudaStream_t stream) {
  auto kern = sweep_knn_kernel<P, W, TM, MINB>;
  const size_t smem = sweep_smem_bytes<P, W>(list_bytes * TM);
  PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int occ = 0;
  PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSweepThreads, smem));
  if (occ < 1) { set_error("kNN sweep variant does not fit on an SM (smem %zu bytes)", smem); return PG_ERR_UNSUPPORTED; }
  const long long n_items = static_cast<long long>(prm.n_rowblocks) * prm.n_splits;
  long long grid = static_cast<long long>(num_sms()) * occ;
  if (grid > n_items) grid = n_items;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), kSweepThreads, smem, stream>>>(prm);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// defined in pg_sweep_knn.cu
int sweep_knn_variant(int variant, int planes, int words, const SweepParams& prm, size_t list_bytes, cudaStream_t s);
int knn_variant_rows_per_cta(int variant);

}  // namespace pg
