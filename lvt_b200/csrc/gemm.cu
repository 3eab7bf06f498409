// lvt_b200 :: bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory, mbarrier pipeline).
//
// One CTA computes one 128 x BN output tile (BN = 128 or 256):
//   warp 0   : TMA producer        (one elected lane)
//   warp 1   : TMEM owner + MMA issuer (one elected lane issues tcgen05.mma)
//   warps 2-5: epilogue, TMEM -> registers -> global (each warp owns the TMEM lane quarter
//              warp_id % 4, i.e. 32 rows of the tile)
// Replaces the cuBLAS calls under torch.bmm / nn.Linear / 1x1x1 Conv3d on the DSFVT path
// (reference: vidgen/modeling/autoregressive/vt_attention.py:63-80,120-128,138 and
// videotransformer.py:57,99,148-156) and their autograd backward.
#include <unordered_map>
#include <mutex>
#include <string>
#include <string.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 192;

struct GemmParams {
  int M, N, K, batch, splits;
  int a_cin, a_zdiv, b_cin, b_zdiv;
  int mode, flags;
  float alpha;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  const float* res;
  const __nv_bfloat16* aux;
  const float* bias;
  int bias_mod;
  int o_cin, o_zdiv;
  long long o_ld, o_s_blk, o_s_zlo, o_s_zhi;
  float* lse;
  const float* delta;
  const float* bank_t;
  const float* bank_h;
  const float* bank_w;
  int bt, bh, bw, heads;
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BANK_OFFSET = BAR_OFFSET + 256;       // barriers + tmem ptr
  static constexpr int TOTAL = BANK_OFFSET + 1024 /*banks*/ + 1024 /*alignment slack*/;
};

LVT_DEVICE_INLINE void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bank = reinterpret_cast<float*>(smem + L::BANK_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int z = blockIdx.z / p.splits;
  const int split = blockIdx.z - z * p.splits;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int kb_begin = split * kb_per;
  const int kb_end = min(kb_total, kb_begin + kb_per);
  const int num_kb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  if (p.mode == LVT_EPI_SOFTMAX && warp >= 2) {
    // stage this head's relative-position banks: [dt | dh | dw]
    const int head = z % p.heads;
    const int nt = 2 * p.bt - 1, nh = 2 * p.bh - 1, nw = 2 * p.bw - 1;
    for (int i = threadIdx.x - 64; i < nt + nh + nw; i += 128) {
      float v;
      if (i < nt) v = p.bank_t[head * nt + i];
      else if (i < nt + nh) v = p.bank_h[head * nh + (i - nt)];
      else v = p.bank_w[head * nw + (i - nt - nh)];
      s_bank[i] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const int a_zlo = z % p.a_zdiv, a_zhi = z / p.a_zdiv;
      const int b_zlo = z % p.b_zdiv, b_zhi = z / p.b_zdiv;
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int k0 = (kb_begin + it) * BK;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* a_dst = smem + s * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        if (!A_MN) {
          tma_load_5d(a_dst, &tm_a, &full_bar[s], k0 % p.a_cin, m0, k0 / p.a_cin, a_zlo, a_zhi);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {
            const int c = m0 + 64 * j;
            tma_load_5d(a_dst + j * (64 * BK * 2), &tm_a, &full_bar[s], c % p.a_cin, k0,
                        c / p.a_cin, a_zlo, a_zhi);
          }
        }
        if (!B_MN) {
          tma_load_5d(b_dst, &tm_b, &full_bar[s], k0 % p.b_cin, n0, k0 / p.b_cin, b_zlo, b_zhi);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int c = n0 + 64 * j;
            tma_load_5d(b_dst + j * (64 * BK * 2), &tm_b, &full_bar[s], c % p.b_cin, k0,
                        c / p.b_cin, b_zlo, b_zhi);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, /*bf16*/ 1, A_MN, B_MN);
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4) {
          // K-major: advance 16 bf16 = 32 B inside the 128 B swizzle row.
          // MN-major: advance 16 k-rows = 2 swizzle atoms of 1024 B.
          const uint64_t adesc = A_MN ? umma_smem_desc(a_base + k4 * 2048, 64 * BK * 2, 1024)
                                      : umma_smem_desc(a_base + k4 * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? umma_smem_desc(b_base + k4 * 2048, 64 * BK * 2, 1024)
                                      : umma_smem_desc(b_base + k4 * 32, 16, 1024);
          umma_bf16_ss(tmem_base, adesc, bdesc, idesc, (it > 0 || k4 > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const long long o_base = (long long)(z / p.o_zdiv) * p.o_s_zhi +
                             (long long)(z % p.o_zdiv) * p.o_s_zlo + (long long)row * p.o_ld;
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    uint32_t r[32];

    if (p.mode == LVT_EPI_LINEAR) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;  // warp-uniform
        if (num_kb > 0) {
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0;
        }
        if (!row_ok) continue;
        const long long off = o_base + (long long)(col0 / p.o_cin) * p.o_s_blk + (col0 % p.o_cin);
        const bool full = (col0 + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
        if (p.bias) {
          const float* bp = p.bias + (p.bias_mod > 0 ? (long long)(row % p.bias_mod) * p.N : 0) + col0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bp + i);
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          } else {
            for (int i = 0; i < 32 && col0 + i < p.N; ++i) v[i] += bp[i];
          }
        }
        if (p.res) {
          const float* rp = p.res + off;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(rp + i);
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          } else {
            for (int i = 0; i < 32 && col0 + i < p.N; ++i) v[i] += rp[i];
          }
        }
        if (p.flags & LVT_GEMM_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (p.flags & LVT_GEMM_MASK) {
          const __nv_bfloat16* ap = p.aux + off;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(ap + i);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
                if (!(lo != 0 && lo < 0x8000u)) v[i + 2 * j] = 0.f;
                if (!(hi != 0 && hi < 0x8000u)) v[i + 2 * j + 1] = 0.f;
              }
            }
          } else {
            for (int i = 0; i < 32 && col0 + i < p.N; ++i)
              if (!(__bfloat162float(ap[i]) > 0.f)) v[i] = 0.f;
          }
        }
        if (p.out_f32) {
          float* op = p.out_f32 + off;
          if (p.flags & LVT_GEMM_ATOMIC) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) red_add_v4(op + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
              for (int i = 0; i < 32 && col0 + i < p.N; ++i) atomicAdd(op + i, v[i]);
            }
          } else if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            for (int i = 0; i < 32 && col0 + i < p.N; ++i) op[i] = v[i];
          }
        }
        if (p.out_bf16) {
          __nv_bfloat16* op = p.out_bf16 + off;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 u;
              u.x = pack_bf16x2(v[i], v[i + 1]);
              u.y = pack_bf16x2(v[i + 2], v[i + 3]);
              u.z = pack_bf16x2(v[i + 4], v[i + 5]);
              u.w = pack_bf16x2(v[i + 6], v[i + 7]);
              *reinterpret_cast<uint4*>(op + i) = u;
            }
          } else {
            for (int i = 0; i < 32 && col0 + i < p.N; ++i) op[i] = __float2bfloat16(v[i]);
          }
        }
      }
    } else if (p.mode == LVT_EPI_SOFTMAX) {
      // One thread owns one query row and all 256 keys of its block (BN == N == 256).
      // v = alpha*acc + B[head, i, j]; causal: j > i -> -1e4 (vt_attention.py:63-74).
      if constexpr (BN == 256) {
        const int nt = 2 * p.bt - 1, nh = 2 * p.bh - 1;
        const float* sb_t = s_bank;
        const float* sb_h = s_bank + nt;
        const float* sb_w = s_bank + nt + nh;
        const int hw = p.bh * p.bw;
        const int ti = row / hw, hi = (row / p.bw) % p.bh, wi = row % p.bw;
        const bool causal = (p.flags & LVT_GEMM_CAUSAL) != 0;
        auto logit = [&](float acc, int j) -> float {
          const int tj = j / hw, hj = (j / p.bw) % p.bh, wj = j % p.bw;
          float v = acc * p.alpha + (sb_t[ti - tj + p.bt - 1] + sb_h[hi - hj + p.bh - 1] +
                                     sb_w[wi - wj + p.bw - 1]);
          if (causal && j > row) v = -1e4f;
          return v;
        };
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, logit(__uint_as_float(r[i]), c0 + i));
        }
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += __expf(logit(__uint_as_float(r[i]), c0 + i) - mx);
        }
        const float inv = 1.f / sum;
        if (row_ok && p.lse) p.lse[(long long)z * p.M + row] = mx + __logf(sum);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
          if (!row_ok) continue;
          __nv_bfloat16* op = p.out_bf16 + o_base + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              e[j] = __expf(logit(__uint_as_float(r[i + j]), c0 + i + j) - mx) * inv;
            uint4 u;
            u.x = pack_bf16x2(e[0], e[1]);
            u.y = pack_bf16x2(e[2], e[3]);
            u.z = pack_bf16x2(e[4], e[5]);
            u.w = pack_bf16x2(e[6], e[7]);
            *reinterpret_cast<uint4*>(op + i) = u;
          }
        }
      }
    } else {  // LVT_EPI_DS: dS = P * (alpha*acc - delta[row])
      const float dl = row_ok ? p.delta[(long long)z * p.M + row] : 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;
        tmem_ld_32x32(taddr + c0, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        const long long off = o_base + (long long)(col0 / p.o_cin) * p.o_s_blk + (col0 % p.o_cin);
        const __nv_bfloat16* ap = p.aux + off;
        __nv_bfloat16* op = p.out_bf16 + off;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const uint4 pu = *reinterpret_cast<const uint4*>(ap + i);
          const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&pu);
          float e[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pf = __bfloat1622float2(pp[j]);
            e[2 * j] = pf.x * (__uint_as_float(r[i + 2 * j]) * p.alpha - dl);
            e[2 * j + 1] = pf.y * (__uint_as_float(r[i + 2 * j + 1]) * p.alpha - dl);
          }
          uint4 u;
          u.x = pack_bf16x2(e[0], e[1]);
          u.y = pack_bf16x2(e[2], e[3]);
          u.z = pack_bf16x2(e[4], e[5]);
          u.w = pack_bf16x2(e[6], e[7]);
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ----------------------------------------------------------------------------------------
// host: TMA descriptor construction (driver entry point fetched at run time: no -lcuda)
// ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

struct MapKey {
  const void* base;
  long long v[12];
  bool operator==(const MapKey& o) const {
    return base == o.base && memcmp(v, o.v, sizeof(v)) == 0;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    for (int i = 0; i < 12; ++i) h = h * 1000003u ^ std::hash<long long>()(k.v[i]);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
std::mutex g_map_mutex;

// Builds the 5-D bf16 tensor map of one operand:
//   dims    (cin, r_extent, n_blocks, zlo, zhi)
//   strides (    ld,        s_blk,    s_zlo, s_zhi)   in elements
//   box     (64, box_rows, 1, 1, 1), SWIZZLE_128B
int make_operand_map(CUtensorMap* out, const void* base, long long c_extent, long long r_extent,
                     int cin, long long ld, long long s_blk, int batch, int zdiv, long long s_zlo,
                     long long s_zhi, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    lvt_set_error("cuTensorMapEncodeTiled driver entry point not available");
    return LVT_ERR_CUDA;
  }
  const long long nblk = (c_extent + cin - 1) / cin;
  const long long zlo = batch < zdiv ? batch : zdiv;
  const long long zhi = (batch + zdiv - 1) / zdiv;
  MapKey key;
  key.base = base;
  const long long kv[12] = {c_extent, r_extent, cin, ld, s_blk, batch, zdiv, s_zlo, s_zhi, box_rows, 0, 0};
  memcpy(key.v, kv, sizeof(kv));
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return LVT_OK;
    }
  }
  cuuint64_t dims[5] = {(cuuint64_t)(cin < c_extent ? cin : c_extent), (cuuint64_t)r_extent,
                        (cuuint64_t)nblk, (cuuint64_t)zlo, (cuuint64_t)zhi};
  auto fix = [](long long s) -> cuuint64_t { return (cuuint64_t)((s > 0 ? s : 8) * 2); };
  cuuint64_t strides[4] = {fix(ld), fix(nblk > 1 ? s_blk : 8), fix(zlo > 1 ? s_zlo : 8),
                           fix(zhi > 1 ? s_zhi : 8)};
  for (int i = 0; i < 4; ++i) {
    if (strides[i] % 16 != 0) {
      lvt_set_error("GEMM operand stride %d (= %llu bytes) is not a multiple of 16 bytes", i,
                    (unsigned long long)strides[i]);
      return LVT_ERR_INVALID;
    }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    lvt_set_error("GEMM operand base pointer is not 16-byte aligned");
    return LVT_ERR_INVALID;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lvt_set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu,%llu "
                  "strides %llu,%llu,%llu,%llu box_rows %d)",
                  (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)dims[2], (unsigned long long)dims[3],
                  (unsigned long long)dims[4], (unsigned long long)strides[0],
                  (unsigned long long)strides[1], (unsigned long long)strides[2],
                  (unsigned long long)strides[3], box_rows);
    return LVT_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
  }
  return LVT_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<grid, NUM_THREADS, L::TOTAL, stream>>>(ta, tb, p);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

template <int BN, int STAGES>
int dispatch_major(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                   bool a_mn, bool b_mn, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, STAGES, false, false>(ta, tb, p, grid, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, STAGES, false, true>(ta, tb, p, grid, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, STAGES, true, false>(ta, tb, p, grid, stream);
  return launch_gemm<BN, STAGES, true, true>(ta, tb, p, grid, stream);
}

}  // namespace

extern "C" int lvt_gemm_bf16(const LvtGemm* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LVT_CHECK_ARG(g != nullptr, "lvt_gemm_bf16: null descriptor");
  LVT_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch > 0, "lvt_gemm_bf16: bad shape %dx%dx%d batch %d",
                g->M, g->N, g->K, g->batch);
  LVT_CHECK_ARG(g->a && g->b, "lvt_gemm_bf16: null operand");
  LVT_CHECK_ARG(g->out_f32 || g->out_bf16, "lvt_gemm_bf16: no output");
  const int splits = g->splits > 0 ? g->splits : 1;
  LVT_CHECK_ARG(splits == 1 || ((g->flags & LVT_GEMM_ATOMIC) && !g->out_bf16 && g->mode == LVT_EPI_LINEAR),
                "lvt_gemm_bf16: split-K needs LVT_GEMM_ATOMIC fp32 output only");
  LVT_CHECK_ARG(!(g->flags & LVT_GEMM_ATOMIC) || (g->out_f32 && !g->res),
                "lvt_gemm_bf16: atomic output needs out_f32 and no residual");
  LVT_CHECK_ARG(g->a_cin > 0 && g->b_cin > 0 && g->o_cin > 0 && g->a_zdiv > 0 && g->b_zdiv > 0 && g->o_zdiv > 0,
                "lvt_gemm_bf16: cin/zdiv must be positive");
  LVT_CHECK_ARG(g->o_cin % 32 == 0 || g->o_cin >= g->N, "lvt_gemm_bf16: o_cin must be a multiple of 32");
  LVT_CHECK_ARG(g->o_ld % 8 == 0 && g->o_s_blk % 8 == 0 && g->o_s_zlo % 8 == 0 && g->o_s_zhi % 8 == 0,
                "lvt_gemm_bf16: output strides must be multiples of 8 elements");
  if (g->flags & LVT_GEMM_MASK) LVT_CHECK_ARG(g->aux_bf16, "lvt_gemm_bf16: MASK needs aux_bf16");

  int bn = 128;
  if (g->mode == LVT_EPI_SOFTMAX) {
    LVT_CHECK_ARG(g->N == 256 && g->out_bf16 && g->bank_t && g->bank_h && g->bank_w &&
                      g->bt * g->bh * g->bw == g->M && g->M == 256 && g->heads > 0,
                  "lvt_gemm_bf16: SOFTMAX mode needs M == N == 256 == bt*bh*bw, banks and out_bf16");
    LVT_CHECK_ARG(2 * (g->bt + g->bh + g->bw) - 3 <= 256, "lvt_gemm_bf16: relative-position banks too large");
    bn = 256;
  } else if (g->mode == LVT_EPI_DS) {
    LVT_CHECK_ARG(g->out_bf16 && g->aux_bf16 && g->delta, "lvt_gemm_bf16: DS mode needs out_bf16, aux_bf16 (P) and delta");
    bn = (g->N % 256 == 0) ? 256 : 128;
  } else {
    LVT_CHECK_ARG(g->mode == LVT_EPI_LINEAR, "lvt_gemm_bf16: unknown epilogue mode %d", g->mode);
    // wide tiles when they still fill the machine
    const long long tiles256 = (long long)((g->M + BM - 1) / BM) * ((g->N + 255) / 256) * g->batch * splits;
    bn = (g->N % 256 == 0 && tiles256 >= 2 * 148) ? 256 : 128;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K; p.batch = g->batch; p.splits = splits;
  p.a_cin = g->a_cin; p.a_zdiv = g->a_zdiv; p.b_cin = g->b_cin; p.b_zdiv = g->b_zdiv;
  p.mode = g->mode; p.flags = g->flags; p.alpha = g->alpha;
  p.out_f32 = g->out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(g->out_bf16);
  p.res = g->res; p.aux = reinterpret_cast<const __nv_bfloat16*>(g->aux_bf16);
  p.bias = g->bias; p.bias_mod = g->bias_mod;
  p.o_cin = g->o_cin; p.o_zdiv = g->o_zdiv;
  p.o_ld = g->o_ld; p.o_s_blk = g->o_s_blk; p.o_s_zlo = g->o_s_zlo; p.o_s_zhi = g->o_s_zhi;
  p.lse = g->lse; p.delta = g->delta;
  p.bank_t = g->bank_t; p.bank_h = g->bank_h; p.bank_w = g->bank_w;
  p.bt = g->bt; p.bh = g->bh; p.bw = g->bw; p.heads = g->heads > 0 ? g->heads : 1;

  CUtensorMap ta, tb;
  int rc;
  // contiguous / strided extents per major-ness
  rc = g->a_mn_major
           ? make_operand_map(&ta, g->a, g->M, g->K, g->a_cin, g->a_ld, g->a_s_blk, g->batch, g->a_zdiv, g->a_s_zlo, g->a_s_zhi, BK)
           : make_operand_map(&ta, g->a, g->K, g->M, g->a_cin, g->a_ld, g->a_s_blk, g->batch, g->a_zdiv, g->a_s_zlo, g->a_s_zhi, BM);
  if (rc) return rc;
  rc = g->b_mn_major
           ? make_operand_map(&tb, g->b, g->N, g->K, g->b_cin, g->b_ld, g->b_s_blk, g->batch, g->b_zdiv, g->b_s_zlo, g->b_s_zhi, BK)
           : make_operand_map(&tb, g->b, g->K, g->N, g->b_cin, g->b_ld, g->b_s_blk, g->batch, g->b_zdiv, g->b_s_zlo, g->b_s_zhi, bn);
  if (rc) return rc;

  dim3 grid((g->N + bn - 1) / bn, (g->M + BM - 1) / BM, g->batch * splits);
  LVT_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "lvt_gemm_bf16: grid too large");
  if (bn == 256) return dispatch_major<256, 4>(ta, tb, p, grid, g->a_mn_major != 0, g->b_mn_major != 0, stream);
  return dispatch_major<128, 3>(ta, tb, p, grid, g->a_mn_major != 0, g->b_mn_major != 0, stream);
}
