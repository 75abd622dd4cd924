// Library plumbing, the token packer and the integer-pipe peak probe.
#include <atomic>

#include "pg_common.cuh"

namespace pg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

cudaError_t temp_alloc(void** ptr, size_t bytes, cudaStream_t s) {
  static bool tuned[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !tuned[dev]) {
    // default release threshold is 0: every free + sync would hand the memory back to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    tuned[dev] = true;
  }
  return cudaMallocAsync(ptr, bytes < 16 ? 16 : bytes, s);
}

// ---- pack: [N][L] tokens -> [rows_padded][planes][words] bit planes ------------------
// One warp per row; lane j holds residue 32w+j, one ballot per plane builds the word.
// Replaces the fp16 staging of prograph.py:726 for integer tokens (HBM-bound: reads the
// tokens once, writes planes*words*4 bytes per row).
template <typename T>
__device__ __forceinline__ bool token_of(T v, int planes, unsigned* tok) {
  long long i = static_cast<long long>(v);
  *tok = static_cast<unsigned>(i);
  return i >= 0 && i < (1ll << planes);
}
template <>
__device__ __forceinline__ bool token_of<float>(float v, int planes, unsigned* tok) {
  const float r = truncf(v);
  *tok = static_cast<unsigned>(static_cast<int>(r));
  return r == v && v >= 0.0f && v < static_cast<float>(1 << planes);
}
template <>
__device__ __forceinline__ bool token_of<double>(double v, int planes, unsigned* tok) {
  const double r = trunc(v);
  *tok = static_cast<unsigned>(static_cast<long long>(r));
  return r == v && v >= 0.0 && v < static_cast<double>(1 << planes);
}
template <>
__device__ __forceinline__ bool token_of<__half>(__half v, int planes, unsigned* tok) {
  return token_of<float>(__half2float(v), planes, tok);
}

template <typename T>
__global__ void pack_kernel(const T* __restrict__ tokens, long long N, int L, long long ld, uint32_t* __restrict__ packed,
                            long long rows_padded, int planes, int words, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long row = warp0; row < rows_padded; row += nwarps) {
    uint32_t* dst = packed + static_cast<size_t>(row) * planes * words;
    for (int w = 0; w < words; ++w) {
      const int l = w * 32 + lane;
      unsigned tok = 0;
      bool ok = true;
      if (row < N && l < L) ok = token_of<T>(tokens[static_cast<size_t>(row) * ld + l], planes, &tok);
      if (!ok) { tok = 0; atomicExch(flag, 1); }
      uint32_t mine = 0;
      for (int p = 0; p < planes; ++p) {
        const uint32_t word = __ballot_sync(0xffffffffu, (tok >> p) & 1u);
        if (lane == p) mine = word;
      }
      if (lane < planes) dst[lane * words + w] = mine;
    }
  }
}

// ---- byte tokens (and raw residue letters) -> bit planes, vectorised ---------------------------
// One thread per (row, word): it loads its 32 one-byte tokens with two 128-bit loads (32-bit or byte
// loads when the row pitch does not allow it), optionally maps letters to tokens through a 256-byte
// table in shared memory (fused tokeniser, prograph.py:454-474: twenty np.where passes producing an
// (N, L) int64 array; byte 0 -- numpy's pad of shorter strings -- and letters outside the alphabet
// map to token 0, as in the reference), and gathers bit p of the four tokens of a 32-bit register
// with one multiply:  ((x & 0x01010101 << p) * (0x01020408 << 4-p)) >> 28  leaves the four bits in the
// top nibble, a funnel shift appends it to the plane word.  3 instructions per 4 tokens and plane
// instead of a byte load and a ballot per token and plane; consecutive threads read consecutive
// 32-byte pieces of a row and write consecutive words of a plane.  HBM-bound: reads L bytes and
// writes planes * words * 4 bytes per row.
struct CharLut { uint8_t t[256]; };

template <int PLANES>
__device__ __forceinline__ void planes_of_32_tokens(const uint32_t (&x)[8], uint32_t (&out)[PLANES]) {
#pragma unroll
  for (int p = 0; p < PLANES; ++p) {
    uint32_t word = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      // tokens 4i .. 4i+3: bit p of every byte -> top nibble of the product
      const uint32_t y = p < 4 ? (x[i] & (0x01010101u << p)) : ((x[i] >> 4) & (0x01010101u << (p - 4)));
      const uint32_t prod = y * (0x01020408u << (4 - (p < 4 ? p : p - 4)));
      word = __funnelshift_l(prod, word, 4);        // word = word << 4 | prod >> 28
    }
    out[p] = word;
  }
}

template <int PLANES, bool CHARS>
__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t* __restrict__ src, long long N, int L, long long ld,
                                                         CharLut lut, uint32_t* __restrict__ packed, long long rows_padded,
                                                         int words, int* __restrict__ flag) {
  __shared__ uint8_t lut_s[256];
  if (CHARS) {
    if (threadIdx.x < 256) lut_s[threadIdx.x] = lut.t[threadIdx.x];
    __syncthreads();
  }
  const long long total = rows_padded * words;
  const bool vec16 = (ld % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  const bool vec4 = (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 4 == 0);
  bool bad = false;
  const int wshift = pow2_shift(words);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long row;
    int w;
    split_index(i, words, wshift, &row, &w);
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 0u;
    const int have = row < N ? min(32, L - w * 32) : 0;      // tokens of this word that exist
    if (have > 0) {
      const uint8_t* at = src + static_cast<size_t>(row) * ld + w * 32;
      if (have == 32 && vec16) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(at));
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(at) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
        x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int left = have - 4 * k;
          if (left >= 4 && vec4) {
            x[k] = __ldcs(reinterpret_cast<const uint32_t*>(at) + k);
          } else if (left > 0) {
            uint32_t v = 0;
            for (int b = 0; b < 4 && b < left; ++b) v |= static_cast<uint32_t>(at[4 * k + b]) << (8 * b);
            x[k] = v;
          }
        }
      }
      if (CHARS) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t v = x[k];
          x[k] = static_cast<uint32_t>(lut_s[v & 0xffu]) | (static_cast<uint32_t>(lut_s[(v >> 8) & 0xffu]) << 8) |
                 (static_cast<uint32_t>(lut_s[(v >> 16) & 0xffu]) << 16) | (static_cast<uint32_t>(lut_s[v >> 24]) << 24);
        }
        // bytes past the end of a short last word were loaded as 0 and must stay 0 whatever lut[0] is
        if (have < 32) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int left = have - 4 * k;
            if (left <= 0) x[k] = 0u;
            else if (left < 4) x[k] &= (1u << (8 * left)) - 1u;
          }
        }
      } else if (PLANES < 8) {
        const uint32_t over = ~(((1u << PLANES) - 1u) * 0x01010101u);
        const uint32_t any = (x[0] | x[1] | x[2]) | (x[3] | x[4] | x[5]) | (x[6] | x[7]);
        if (any & over) {           // a token does not fit the planes: flag it and pack it as 0
          bad = true;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t keep = 0;
            for (int b = 0; b < 4; ++b)
              if ((((x[k] >> (8 * b)) & 0xffu) >> PLANES) == 0u) keep |= 0xffu << (8 * b);
            x[k] &= keep;
          }
        }
      }
    }
    uint32_t out[PLANES];
    planes_of_32_tokens<PLANES>(x, out);
    uint32_t* dst = packed + static_cast<size_t>(row) * PLANES * words + w;
#pragma unroll
    for (int p = 0; p < PLANES; ++p) dst[p * words] = out[p];
  }
  if (bad && flag != nullptr) atomicExch(flag, 1);
}

// ---- integer pipe peak probe ------------------------------------------------------------
// Register-only loops with the instruction mix of the Hamming inner loop; the achieved
// lane-op rate is the "speed of light" the sweep kernel's roofline fraction is quoted on.
__device__ __forceinline__ unsigned probe_mad(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

template <int MIX>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, uint32_t seed, unsigned one) {
  uint32_t a[8], b[5];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < 5; ++i) b[i] = seed + blockIdx.x * 977u + i;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MIX == 0) {
      // per "word": 5 LOP3 + 1 POPC + 1 add (as IMAD on the FMA pipe, exactly like the sweep
      // kernel's inner loop), 8 independent words like W=8
      uint32_t m[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        m[w] = a[w] ^ b[0];
        m[w] |= a[(w + 1) & 7] ^ b[1];
        m[w] |= a[(w + 2) & 7] ^ b[2];
        m[w] |= a[(w + 3) & 7] ^ b[3];
        m[w] |= a[(w + 4) & 7] ^ b[4];
      }
#pragma unroll
      for (int w = 0; w < 8; ++w) acc = probe_mad(__popc(m[w]), one, acc);
      if ((it & 15) == 15) {
#pragma unroll
        for (int i = 0; i < 5; ++i) b[i] = b[i] * 3u + acc;  // keep the stream operand changing
      }
    } else if (MIX == 1) {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        a[w] = (a[w] ^ b[0]) | (a[(w + 1) & 7] & b[1]);
        a[w] = (a[w] ^ b[2]) | (a[(w + 3) & 7] & b[3]);
        a[w] = (a[w] ^ b[4]) | (a[(w + 5) & 7] & b[0]);
        a[w] = (a[w] ^ b[1]) | (a[(w + 7) & 7] & b[2]);
        a[w] = (a[w] ^ b[3]) | (a[(w + 2) & 7] & b[4]);
      }
    } else {
#pragma unroll
      for (int w = 0; w < 8; ++w) a[w] = __popc(a[w]) + b[w % 5];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a[i];
  if (acc == 0x12345678u) out[0] = acc;  // never true in practice; keeps the loops alive
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_version(void) { return 100; }
const char* pg_last_error(void) { return g_err; }

int pg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return PG_OK;
}

int64_t pg_launch_count(int reset) {
  const long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int pg_packed_words(int L) {
  if (L <= 0) return 0;
  const int w = (L + 31) / 32;
  if (w <= 1) return 1;
  if (w <= 2) return 2;
  if (w <= 4) return 4;
  return (w + 7) / 8 * 8;
}
int64_t pg_packed_rows(int64_t N) { return N <= 0 ? 0 : round_up(N, kStreamRowPad); }
size_t pg_packed_bytes(int64_t N, int L, int planes) {
  return static_cast<size_t>(pg_packed_rows(N)) * planes * pg_packed_words(L) * 4;
}

int pg_pack_tokens(const void* tokens, int dtype, int64_t N, int L, int64_t ld, uint32_t* packed, int planes, int words,
                   int* flag, void* stream) {
  PG_CHECK_ARG(tokens && packed && flag, "null pointer");
  PG_CHECK_ARG(N > 0 && L > 0 && ld >= L, "bad shape N=%lld L=%d ld=%lld", (long long)N, L, (long long)ld);
  PG_CHECK_ARG(planes >= 1 && planes <= 16, "planes must be in [1,16]");
  PG_CHECK_ARG(words * 32 >= L, "words=%d too small for L=%d", words, L);
  const long long rows_padded = pg_packed_rows(N);
  const int threads = 256;
  long long blocks = ceil_div(rows_padded * 32, threads);
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == PG_U8 && (planes == 5 || planes == 8)) {
    // the format tokens have on the device: vectorised kernel, one thread per (row, word)
    long long vb = ceil_div(rows_padded * words, threads);
    if (vb > cap) vb = cap;
    const CharLut none = {};
    if (planes == 5)
      pack_bytes_kernel<5, false><<<static_cast<unsigned>(vb), threads, 0, s>>>(
          static_cast<const uint8_t*>(tokens), N, L, ld, none, packed, rows_padded, words, flag);
    else
      pack_bytes_kernel<8, false><<<static_cast<unsigned>(vb), threads, 0, s>>>(
          static_cast<const uint8_t*>(tokens), N, L, ld, none, packed, rows_padded, words, flag);
    PG_LAUNCH_CHECK();
    return PG_OK;
  }
#define PG_PACK(T)                                                                                         \
  pack_kernel<T><<<static_cast<unsigned>(blocks), threads, 0, s>>>(static_cast<const T*>(tokens), N, L, ld, packed, \
                                                                    rows_padded, planes, words, flag)
  switch (dtype) {
    case PG_U8: PG_PACK(uint8_t); break;
    case PG_I16: PG_PACK(int16_t); break;
    case PG_I32: PG_PACK(int32_t); break;
    case PG_I64: PG_PACK(long long); break;
    case PG_F16: PG_PACK(__half); break;
    case PG_F32: PG_PACK(float); break;
    case PG_F64: PG_PACK(double); break;
    default: set_error("unsupported token dtype %d", dtype); return PG_ERR_INVALID;
  }
#undef PG_PACK
  PG_LAUNCH_CHECK();
  return PG_OK;
}

int pg_pack_chars(const uint8_t* chars, int64_t N, int L, int64_t ld, const uint8_t* lut256_host, uint32_t* packed,
                  int planes, int words, void* stream) {
  PG_CHECK_ARG(chars && lut256_host && packed, "null pointer");
  PG_CHECK_ARG(N > 0 && L > 0 && ld >= L, "bad shape N=%lld L=%d ld=%lld", (long long)N, L, (long long)ld);
  PG_CHECK_ARG((planes == 5 || planes == 8) && words * 32 >= L, "bad planes / words");
  CharLut lut;
  for (int i = 0; i < 256; ++i) {
    PG_CHECK_ARG(lut256_host[i] < (1u << planes), "token %d of letter %d does not fit %d planes", lut256_host[i], i, planes);
    lut.t[i] = lut256_host[i];
  }
  const long long rows_padded = pg_packed_rows(N);
  const int threads = 256;
  long long blocks = ceil_div(rows_padded * words, threads);
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (planes == 5)
    pack_bytes_kernel<5, true><<<static_cast<unsigned>(blocks), threads, 0, s>>>(chars, N, L, ld, lut, packed, rows_padded,
                                                                                words, nullptr);
  else if (planes == 8)
    pack_bytes_kernel<8, true><<<static_cast<unsigned>(blocks), threads, 0, s>>>(chars, N, L, ld, lut, packed, rows_padded,
                                                                                words, nullptr);
  else {
    set_error("pg_pack_chars supports 5 or 8 planes, got %d", planes);
    return PG_ERR_UNSUPPORTED;
  }
  PG_LAUNCH_CHECK();
  return PG_OK;
}

int pg_measure_int_peak(int mix, int iters, double* lane_ops_per_s, double* ms_out) {
  PG_CHECK_ARG(mix >= 0 && mix <= 2 && iters > 0, "bad probe arguments");
  uint32_t* out = nullptr;
  PG_CUDA(cudaMalloc(&out, 64));
  const int blocks = num_sms() * 8, threads = 256;
  cudaEvent_t a, b;
  PG_CUDA(cudaEventCreate(&a));
  PG_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    PG_CUDA(cudaEventRecord(a, 0));
    if (mix == 0) int_peak_kernel<0><<<blocks, threads>>>(out, iters, 12345u + rep, 1u);
    else if (mix == 1) int_peak_kernel<1><<<blocks, threads>>>(out, iters, 12345u + rep, 1u);
    else int_peak_kernel<2><<<blocks, threads>>>(out, iters, 12345u + rep, 1u);
    PG_CUDA(cudaEventRecord(b, 0));
    PG_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    PG_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  PG_CUDA(cudaGetLastError());
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  // lane-ops counted per iteration and thread: mix0 8*(5+1+1)=56, mix1 8*5*2=80 LOP3, mix2 8 POPC + 8 IADD
  const double per_iter = mix == 0 ? 56.0 : (mix == 1 ? 80.0 : 8.0);
  const double total = per_iter * iters * static_cast<double>(blocks) * threads;
  if (lane_ops_per_s) *lane_ops_per_s = total / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return PG_OK;
}

}  // extern "C"
