// Fused flash-style self attention for the 3D token sequence (MONAI SABlock core as used by vit.py:438-443):
//     out[b, :, h*64:(h+1)*64] = softmax(Q_h K_h^T / sqrt(64)) V_h,   S = 2049 tokens, 12 heads x 64.
// The reference materialises the [B,12,2049,2049] score tensor (201 MB/layer/volume in fp32); here scores never
// leave the SM: S = Q K^T accumulates in TMEM, the softmax warps turn it into bf16 P (also TMEM), and O += P V runs
// with P as the TMEM-resident A operand.
//
// CTA = one 128-query tile of one (batch, head); 2 CTAs are co-resident per SM (256 TMEM columns each).
// Keys are processed in 64-key steps with DOUBLE-BUFFERED score and probability tiles:
//     TMEM columns   0.. 63  S0      64..127  S1     128..159  P0     160..191  P1     192..255  O
//   warp 0 lane 0  TMA producer: Q once, then K / V 128-key tiles straight out of the fused qkv activation
//                  [B*S, 2304] (column offsets 0 / 768 / 1536 + h*64) -- no head-split copies are ever made;
//   warp 1 lane 0, warp 0 lane 1   MMA issuers of the even / odd steps: woken by p_full(t) a thread issues
//                  O += P[b] V (TS: P from TMEM, V is the MN-major B operand straight from its row-major tile) and then
//                  S[b] = Q K^T of step t+2 (SS, both K-major, N = 64), so that s_full(t+2) also certifies P[b] is free;
//   warps 2..5     single-pass online softmax in fp32 (exp2 domain, lazy rescale): ONE mbarrier wait per step
//                  (s_full), 64 scores per thread in registers; O correction only when the running max moved by more
//                  than 2^8; final normalise + bf16 store.
// Why this shape (measured on B200: tools/umma_latency.cu, umma_throughput.cu, attn_trace.py, profiles/README.md): a
// barrier hop costs 200-400 cycles even when the phase is already complete, four tcgen05.mma issues + commit ~440 cycles
// of the issuing thread on a loaded SM, and the step time is the cycle softmax(t) -> issuer -> softmax(t+2).  Splitting
// the issuing work by MMA group (Q K^T thread / P V thread) put four hops on that cycle, a single issuing thread was
// itself the bottleneck (~1200 busy cycles per step); splitting by step parity has two hops and half the work per thread.
// Keys past the end of the sequence (2049 = 32*64 + 1) are masked in the last step; softmax warps whose 32 query rows
// are all past the end (last query tile) skip the exponentials.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <type_traits>

namespace hs {

extern void count_launch();

namespace {

constexpr int QT = 128;                    // queries per CTA
constexpr int KT = 128;                    // keys per TMA tile
constexpr int KS = 64;                     // keys per softmax / MMA step
#ifndef HSENET_ATT_K_STAGES
#define HSENET_ATT_K_STAGES 3
#endif
#ifndef HSENET_ATT_V_STAGES
#define HSENET_ATT_V_STAGES 2
#endif
constexpr int K_STAGES = HSENET_ATT_K_STAGES;
constexpr int V_STAGES = HSENET_ATT_V_STAGES;
constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int SUB_BYTES = KS * 128;        // 64 rows x 128 B
constexpr int ATT_THREADS = 224;   // producer warp, two MMA-issuing warps, four softmax warps
#ifndef HSENET_ATT_TRACE_WARP
#define HSENET_ATT_TRACE_WARP 3
#endif
#ifndef HSENET_ATT_SINGLE_ISSUER
#define HSENET_ATT_SINGLE_ISSUER 1
#endif
constexpr bool kSingleIssuer = HSENET_ATT_SINGLE_ISSUER != 0;
#ifndef HSENET_ATT_PV_INTERLEAVE
#define HSENET_ATT_PV_INTERLEAVE 1
#endif
constexpr bool kPvInterleave = HSENET_ATT_PV_INTERLEAVE != 0;
// (Tried and removed: publishing the P store of step t only after the score load of step t+1 has been issued -- +4 %: the
// p_full -> P V -> Q K^T -> s_full chain is on the cycle.)
#ifndef HSENET_ATT_PARK
#define HSENET_ATT_PARK 1
#endif
// (Tried and removed: probing the NEXT step's s_full early -- result in a register: +4..9 %; deferred into a named PTX predicate
// that is only read one step later: +3..7 %.)
// waits of the producer / issuer warps (1 = parked try_wait with a suspend-time hint, 2 = also the softmax warps)
__device__ __forceinline__ void ctl_wait(uint64_t* bar, uint32_t parity) {
  if (HSENET_ATT_PARK >= 1) mbar_wait_parked(bar, parity); else mbar_wait_nocall(bar, parity);
}
__device__ __forceinline__ void smx_wait(uint64_t* bar, uint32_t parity) {
  if (HSENET_ATT_PARK >= 2) mbar_wait_parked(bar, parity); else mbar_wait_nocall(bar, parity);
}
constexpr int kDefaultPoly = 2;      // measured: 2/8 -> -2.6 %, 4/8 -> +5 % (profiles/README.md)
constexpr int TMEM_COLS = 256;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
constexpr int ATT_SMEM = (1 + K_STAGES + V_STAGES) * TILE_BYTES + 1024 + 256;

struct AttBarriers {
  uint64_t q_full;
  uint64_t k_full[K_STAGES], k_empty[K_STAGES];
  uint64_t v_full[V_STAGES], v_empty[V_STAGES];
  uint64_t s_full[3], p_full[3], pv_done[2];
  uint32_t tmem_base;
};

// Trace build only (HSENET_NVCC_EXTRA=-DHSENET_ATT_TRACE, tools/attn_trace.py): per-phase clock64 sums of one softmax
// warp and of the two issuing threads of CTA 0.  A clock read does not wait for earlier asynchronous instructions to
// COMPLETE, only to issue: a slot also collects stalls caused by what was issued just before it.
#ifdef HSENET_ATT_TRACE
__device__ unsigned long long g_att_trace[48];   // [0,16) softmax warp 2, [16,32) Q K^T issuer, [32,48) P V issuer
#define ATT_TR(i)                        \
  do {                                   \
    const long long _n = clock64();      \
    tr[i] += _n - tlast;                 \
    tlast = _n;                          \
  } while (0)
#define ATT_TR_DECL long long tr[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tlast = clock64(); const long long tstart = tlast
#define ATT_TR_DUMP(base)                                                              \
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {                         \
    for (int i = 0; i < 10; ++i) g_att_trace[(base) + i] = tr[i];                      \
    g_att_trace[(base) + 14] = clock64() - tstart;                                     \
    g_att_trace[(base) + 15] = nsub;                                                   \
  }
__device__ long long g_att_times[14][48];         // rows 0..7: p_full arrive of softmax warp r; 8: s_full seen (warp 3); 9: issuer woken; 10: issue end; 11: issuer has its K/V operands; 12: top of the issuer step; 13: after the v_full wait
#define ATT_TS(row, step)                                                                          \
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (step) < 48) g_att_times[row][step] = clock64()
#define ATT_TV(row, step, val)                                                                     \
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_att_times[row][step] = (val)
#else
#define ATT_TR(i)
#define ATT_TR_DECL
#define ATT_TR_DUMP(base)
#define ATT_TS(row, step)
#define ATT_TV(row, step, val)
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^y for two values on the FMA / ALU pipes (no MUFU): y = n + f with n = round(y), f in [-0.5, 0.5]; cubic for 2^f
// (relative error < 7e-4, inside bf16's 2^-9 rounding); the exponent is patched in with integer adds.  Used for a
// fraction of the exponentials: the MUFU pipe (16 ex2/clk/SM) is what bounds the softmax.
__device__ __forceinline__ void exp2_poly2(float y0, float y1, float& e0, float& e1) {
  y0 = fmaxf(y0, -126.0f);
  y1 = fmaxf(y1, -126.0f);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);        // 1.5 * 2^23
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t y = pack2(y0, y1);
  const uint64_t t = add2(y, magic);                             // integer part lands in the low mantissa bits
  const uint64_t n = add2(t, nmagic);
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), y);            // f = y - n
  uint64_t p = fma2(f, pack2(0.0555041f, 0.0555041f), pack2(0.2402265f, 0.2402265f));
  p = fma2(p, f, pack2(0.6931472f, 0.6931472f));
  p = fma2(p, f, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// POLY = how many of every 8 exponentials run on the FMA pipe instead of MUFU (0, 2 or 4)
// NBUF = number of score buffers in flight.  NBUF = 2: the round-1 layout (S0|S1|P0|P1|O).  NBUF = 3 (default since
// round 2): three 64-column score buffers with P ALIASED onto the first 32 columns of its own scores (a thread
// overwrites only columns it has already read into registers; Q K^T (t+3) is issued after P V (t) by the same thread, so
// the MMA pipe orders the overwrite):    0..63 S0/P0   64..127 S1/P1   128..191 S2/P2   192..255 O.
// Why three: tools/attn_trace.py on the two-buffer kernel shows the per-buffer cycle  softmax(t) [~1100 clk] -> p_full
// hop [~250] -> 4 P V issues + commits [~500] -> 4 Q K^T (t+2) issues + commits [~500] -> s_full hop [~250]  = ~2650 clk
// per TWO steps, i.e. the softmax warps are starved ~30 % of the time by the serial issue path of their own buffer.
// With a third buffer the same cycle has three steps of softmax time to hide in.
template <int POLY, int NBUF>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out,
                 int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = sK + K_STAGES * TILE_BYTES;
  AttBarriers* bars = reinterpret_cast<AttBarriers*>(sV + V_STAGES * TILE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = b * S;                      // first row of this volume in the [B*S, 2304] activation
  const int ntiles = (S + KT - 1) / KT;        // 128-key TMA tiles
  const int nsub = (S + KS - 1) / KS;          // 64-key steps

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < K_STAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 2); }   // one commit per issuer
    for (int s = 0; s < V_STAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 2); }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 4);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&bars->pv_done[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
  pdl_prologue_done();      // everything above is independent of the previous kernel's output

  // ===================== MMA issuer of the steps t = par, par+2, ... (two threads: par = 0, 1) =====================
  // Woken by p_full(t), a thread issues P V (t) and then Q K^T (t+2) back to back.  Splitting the issuing work by step
  // parity instead of by MMA group keeps two barrier hops on the cycle softmax(t) -> ... -> softmax(t+2) (a Q K^T / P V
  // role split has four: p_full, pv_done -> Q K^T issuer, s_full; profiles/README.md) while each thread is busy only
  // every other step (one thread for everything is itself the bottleneck at ~1200 busy cycles per step).  p_full(t)
  // implies that softmax(t) holds S[par] in registers (no s_free barrier), and since one thread's MMAs complete in order
  // s_full(t+2) implies that P V (t) has consumed P[par].  K / V stages are released by one commit from EACH thread, and
  // the P V groups of the two threads are kept in step order (deterministic accumulation into O).
  // TMEM column of score buffer b / of the packed probabilities of buffer b
  auto col_s = [](int b) -> uint32_t { return static_cast<uint32_t>(b * KS); };
  auto col_p = [](int b) -> uint32_t { return NBUF == 3 ? static_cast<uint32_t>(b * KS) : COL_P + static_cast<uint32_t>(b * 32); };
  auto mma_issuer = [&](const int par) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(QT, KS, 0, 0);          // S[128q x 64k] = Q K^T
    constexpr uint32_t idesc_pv = make_idesc_bf16(QT, kHeadDim, 0, 1);    // O[128q x 64d] += P V (V MN-major)
    const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
    auto issue_qk = [&](const int t) {                                    // k_full of its tile already waited for
      const int ks = (t >> 1) % K_STAGES, b = t % NBUF;
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + ks * TILE_BYTES + (t & 1) * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + col_s(b), qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        tc_commit(&bars->s_full[b]);
        tc_commit(&bars->k_empty[ks]);
      }
      __syncwarp();
    };
    // Step t is served by thread (t & 1): P V (t), then Q K^T (t + NBUF) into the buffer P V (t) has just read.
    // Prologue: Q K^T (s), s < NBUF, is issued by the thread that would have issued it in the loop, thread ((s - NBUF) & 1),
    // so that every 128-key K tile is released by exactly one commit from each thread.
    ctl_wait(&bars->q_full, 0);
    for (int s0 = 0; s0 < NBUF && s0 < nsub; ++s0) {
      if (((s0 + NBUF) & 1) != par) continue;
      ctl_wait(&bars->k_full[(s0 >> 1) % K_STAGES], 0);
      issue_qk(s0);
    }
    ATT_TR_DECL;
    for (int t = par; t < nsub; t += 2) {
      const int j = t >> 1, vs = j % V_STAGES, t2 = t + NBUF, b = t % NBUF;
      // operands of this iteration's MMAs first: long satisfied, kept off the critical path
      ctl_wait(&bars->v_full[vs], (j / V_STAGES) & 1);
      if (t2 < nsub) ctl_wait(&bars->k_full[(t2 >> 1) % K_STAGES], ((t2 >> 1) / K_STAGES) & 1);
      // P V (t-1), issued by the other thread, must have retired before P V (t) is issued: both accumulate into the same
      // fp32 tile, and an occasional swap of two accumulations would make the last bits differ from run to run (and
      // P V (0) initialises O).  It retires long before p_full(t) arrives, so this wait is off the critical path too.
      if (t >= 1) ctl_wait(&bars->pv_done[par ^ 1], ((t - 1) >> 1) & 1);
      ATT_TR(0);
      ctl_wait(&bars->p_full[b], (t / NBUF) & 1);          // softmax t done: P[b] stored, S[b] in registers
      tc_fence_after();
      ATT_TR(1);
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES + par * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KS / 16; ++k) {
          // A: 16 keys = 8 TMEM columns of packed bf16 pairs; B: 16 key rows = 2048 bytes = +128 (16-byte units)
          umma_ts(tmem_base + COL_O, tmem_base + col_p(b) + 8 * k, vdesc + 128 * k, idesc_pv, (t | k) != 0 ? 1u : 0u);
        }
        tc_commit(&bars->pv_done[par]);
        tc_commit(&bars->v_empty[vs]);
      }
      __syncwarp();
      ATT_TR(2);
      if (t2 < nsub) issue_qk(t2);
      ATT_TR(3);
    }
    ATT_TR_DUMP(16 + 16 * par);
  };

  // Roles are whole WARPS that stay convergent; the single-thread instructions (TMA, tcgen05.mma, tcgen05.commit) sit
  // inside elect_one() regions.  Round 1 ran each role as one lane of a diverged warp ("if (lane == 0)"): ptxas then
  // wraps every instruction that takes uniform-register operands (UTCHMMA, UTCBAR, UTMALDG, SYNCS) in a per-lane
  // ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- ~100 issue cycles per 32-cycle MMA (cuobjdump of the round-1 kernel).
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, &bars->q_full, h * kHeadDim, row0 + q0);
    }
    __syncwarp();
    // K runs one tile ahead of V: Q K^T is issued up to three steps before the P V of the same keys, and a K stage
    // is released that much earlier than the V stage of the same tile -- loading K(j+1) only after V(j) had found
    // its stage free (the round-1 order) left the K tile two steps of lead, less than a TMA round trip under load
    auto load_k = [&](const int j) {
      const int ks = j % K_STAGES;
      ctl_wait(&bars->k_empty[ks], ((j / K_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->k_full[ks], TILE_BYTES);
        tma_load_2d_hint(sK + ks * TILE_BYTES, &tmQKV, &bars->k_full[ks], kHidden + h * kHeadDim, row0 + j * KT,
                         kEvictLast);
      }
      __syncwarp();
    };
    load_k(0);
    for (int j = 0; j < ntiles; ++j) {
      const int vs = j % V_STAGES;
      if (j + 1 < ntiles) load_k(j + 1);
      ctl_wait(&bars->v_empty[vs], ((j / V_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->v_full[vs], TILE_BYTES);
        tma_load_2d_hint(sV + vs * TILE_BYTES, &tmQKV, &bars->v_full[vs], 2 * kHidden + h * kHeadDim,
                         row0 + j * KT, kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mma_issuer(0);
  } else if (warp == 2) {
    mma_issuer(1);
  } else {
    // ===================== softmax / correction / epilogue warps =====================
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int qi = q0 + quarter * 32 + lane;             // query index inside the volume
    const bool warp_live = (q0 + quarter * 32) < S;      // warp-uniform: any valid query row in this warp?
    const float c = 0.125f * 1.4426950408889634f;       // head_dim^-0.5 * log2(e)
    float m = -INFINITY;                                 // running reference max (log2 domain, already scaled)
    float l = 0.f;
    ATT_TR_DECL;
    // one 64-key step; MASKED (compile time) only for a last step that runs past the end of the sequence -- kept out
    // of the main loop on purpose: left as a run-time test the compiler turns the 64 per-key checks into selects that
    // execute on EVERY step (195 of ~530 instructions per step in the first version)
    auto softmax_step = [&](const int t, auto masked) {
      const int bsel = t % NBUF;
      // S[bsel] ready and P[bsel] free
      smx_wait(&bars->s_full[bsel], (t / NBUF) & 1);
      tc_fence_after();
      ATT_TR(0);
      uint32_t x[64];
      uint32_t pk[32];
      float alpha = 1.f;
      if (warp_live) {
        tmem_ld32(tmem_base + lane_base + col_s(bsel), *reinterpret_cast<uint32_t(*)[32]>(&x[0]));
        tmem_ld32(tmem_base + lane_base + col_s(bsel) + 32, *reinterpret_cast<uint32_t(*)[32]>(&x[32]));
      }
      if (warp_live) tmem_ld_wait();
      ATT_TR(1);
      ATT_TR(2);
      if (warp_live) {
        const int kbase = t * KS;
        if constexpr (decltype(masked)::value) {         // last step: mask keys past the sequence end
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (kbase + i >= S) x[i] = 0xff800000u;      // -inf
        }
        float mx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = __uint_as_float(x[u]);
#pragma unroll
        for (int i = 4; i < 64; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(x[i + u]));
        }
        const float tm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c;
        // running max with lazy rescale: only move the reference max when it grows by more than 2^8
        if (tm > m + 8.0f) {
          alpha = ex2(m - tm);                           // m = -inf on the first step -> 0
          m = tm;
        }
        ATT_TR(3);
        // scale-subtract and row sum on packed fp32x2 (FFMA2 / FADD2: two lanes per issued instruction)
        const uint64_t c2 = pack2(c, c), nm2 = pack2(-m, -m);
        uint64_t rs2 = pack2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          float e[8];
#pragma unroll
          for (int u = 0; u < 8; u += 2)
            unpack2(fma2(pack2(__uint_as_float(x[i + u]), __uint_as_float(x[i + u + 1])), c2, nm2), e[u], e[u + 1]);
#pragma unroll
          for (int u = 0; u < 8 - POLY; ++u) e[u] = ex2(e[u]);                                        // MUFU pipe
#pragma unroll
          for (int u = 8 - POLY; u < 8; u += 2) exp2_poly2(e[u], e[u + 1], e[u], e[u + 1]);            // FMA pipe
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            rs2 = add2(rs2, pack2(e[u], e[u + 1]));
            pk[(i + u) >> 1] = pack_bf16x2(e[u], e[u + 1]);
          }
        }
        float rs0, rs1;
        unpack2(rs2, rs0, rs1);
        l = l * alpha + (rs0 + rs1);
        ATT_TR(4);
        tmem_st32(tmem_base + lane_base + col_p(bsel), pk);
        ATT_TR(5);
      }
      // O correction (rare after the first steps): P V (t-1) must have retired before O is rescaled
      if (t > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        smx_wait(&bars->pv_done[(t & 1) ^ 1], ((t - 1) >> 1) & 1);
        if (t >= 2) smx_wait(&bars->pv_done[t & 1], ((t >> 1) - 1) & 1);   // the other issuing thread's P V (t-2)
        tc_fence_after();
        uint32_t o[32];
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          tmem_ld32(tmem_base + lane_base + COL_O + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_base + lane_base + COL_O + ch * 32, o);
        }
      }
      ATT_TR(6);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[bsel]);
      ATT_TR(7);
    };
    const bool ragged = (S % KS) != 0;
    for (int t = 0; t < nsub - (ragged ? 1 : 0); ++t) softmax_step(t, std::false_type{});
    if (ragged) softmax_step(nsub - 1, std::true_type{});
    if (warp == HSENET_ATT_TRACE_WARP && lane == 0) { ATT_TR_DUMP(0); }
    // ---- epilogue: O / l -> bf16 -> out[b*S + qi, h*64 .. h*64+63] -----------------------------------------------
    if (nsub >= 2) smx_wait(&bars->pv_done[(nsub - 2) & 1], ((nsub - 2) >> 1) & 1);
    smx_wait(&bars->pv_done[(nsub - 1) & 1], ((nsub - 1) >> 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    // training forward: log2-domain log-sum-exp of the scaled scores, lse = m + log2(l), so that the backward kernels
    // rebuild P = 2^(s c - lse) without a second pass; rows past the sequence end get +inf (their P is then exactly 0)
    if (lse_out != nullptr) {
      const int sp = ((S + QT - 1) / QT) * QT;
      lse_out[(static_cast<long>(b) * kHeads + h) * sp + qi] = qi < S ? m + log2f(l) : INFINITY;
    }
    uint32_t o[32];
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld32(tmem_base + lane_base + COL_O + ch * 32, o);
      tmem_ld_wait();
      if (qi < S) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long>(row0) + qi) * kHidden + h * kHeadDim + ch * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          dst[g] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// Largest squared key norm per (volume, head): kmax2[b*12 + h] = max_k |K[b,k,h,:]|^2 (non-negative floats compare like
// their bit patterns, so an integer atomicMax is exact and order-independent).  Feeds the max-free softmax below.
__global__ void __launch_bounds__(192) attn_kmax_kernel(const __nv_bfloat16* __restrict__ qkv, int* __restrict__ kmax_bits,
                                                        int S) {
  __shared__ float red[16][kHeads];
  const int h = threadIdx.x, ry = threadIdx.y, b = blockIdx.y;
  const int r = blockIdx.x * 16 + ry;
  float n2 = 0.f;
  if (r < S) {
    const uint4* p = reinterpret_cast<const uint4*>(qkv + (static_cast<long>(b) * S + r) * (3 * kHidden) + kHidden + h * kHeadDim);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = __ldg(p + j);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
        n2 = fmaf(f.x, f.x, fmaf(f.y, f.y, n2));
      }
    }
  }
  red[ry][h] = n2;
  __syncthreads();
  if (ry == 0) {
    float m = red[0][h];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, red[i][h]);
    atomicMax(kmax_bits + b * kHeads + h, __float_as_int(m));
  }
}

// =====================================================================================================================
// Key-split variant (default since round 2): EIGHT softmax warps per CTA, two per TMEM lane quarter.  The two warps of a
// quarter share the same 32 query rows and split every 64-key step by KEY half (keys 0..31 / 32..63): each keeps its own
// running max, row sum and its OWN fp32 output accumulator (O_A / O_B in TMEM, fed by the K = 32 halves of P V), so no
// row statistic is ever exchanged inside the loop -- the two partial softmaxes are merged once, in the epilogue
// (m* = max(m_A, m_B); O = (O_A 2^(m_A-m*) + O_B 2^(m_B-m*)) / (l_A 2^(m_A-m*) + l_B 2^(m_B-m*))).
// Why: the round-1 kernel was a per-warp LATENCY chain (mbarrier probe ~230 cycles, tcgen05.ld, max, 64 exponentials,
// tcgen05.st + wait ~400, arrive) with one softmax warp per scheduler and CTA -- a CTA ran as fast alone on an SM as with
// a partner and no unit was above 50 % (profiles/README.md).  Eight warps halve the serial exp/max work per warp-step
// and put four independent chains on every scheduler (2 CTAs x 2 warps), without the shared-memory max exchange that
// made the row-sharing v3/v9 slower.  P is aliased onto the first half of its own score columns (a warp overwrites only
// columns it has already read into registers; Q K^T (t+2) is issued after P V (t) by the same thread, so the MMA pipe
// orders the overwrite), which keeps the CTA at 256 TMEM columns and two CTAs per SM:
//     TMEM columns   0.. 63  S0 (P0_A at 0..15, P0_B at 32..47)    64..127  S1 (P1_A, P1_B)    128..191 O_A    192..255 O_B
// Issuing threads, TMA rings and the step-parity protocol are those of the kernel above; p_full counts 8 warps.
constexpr int ATS_THREADS = 384;   // four control warps (one per scheduler: producer / issuer by SM slot), eight softmax warps
// The two CTAs that share an SM put their control warps on DIFFERENT schedulers (slot 0: producer warp 0, issuer warp 1;
// slot 1: producer warp 2, issuer warp 3).  tools/attn_timeline.py: the two softmax warps that share a scheduler with
// their CTA's issuing warp arrive 600-700 cycles after the other six at EVERY step (the lag follows the issuer when its
// warp is moved), and p_full needs the slowest warp: the issue burst (8 tcgen05.mma + 4 commits, 350 cycles, held by the
// tensor pipe's queue) starts exactly when those two warps begin their next step.  Spreading the two CTAs' issuers over
// two schedulers does not remove that lag (each CTA still suffers from its own issuer) but measured +3 % at batch 32.
#ifndef HSENET_ATT_OBSERVE_PV
#define HSENET_ATT_OBSERVE_PV 1
#endif
#ifndef HSENET_ATT_SPREAD
#define HSENET_ATT_SPREAD 1
#endif
// (Tried and removed: halving the commits per step -- K / V stages released once per 128-key tile, P V completion inferred
// from s_full(t+1) -- made the kernel 3 % SLOWER and the issue burst longer, 350 -> 590 cycles: the issuing thread is held
// by the tensor pipe's short queue, eight back-to-back MMAs block it longer than 4 + 4 with commits in between.)
// (Tried and removed: computing the single live row of the 17th query tile (S = 2049 = 16 x 128 + 1, the cls row) on the
// CUDA cores instead of running a whole tile CTA for it.  A tile CTA only lives ~27 us (147 us x 296 slots / 1632 CTAs); the
// row path -- 2049 dot products, softmax, P V from L2 with 384 threads -- is bound by ~17 dependent L2 round trips and took
// longer: 146.9 -> 154.8 us at batch 8, equal at batch 32.  S = 2048 measures 136 us: the 17th tile costs 8 %.)
// (Tried and removed: folding the orphan KEY (S % 64 == 1 costs every CTA a 33rd, masked step: 2.9 % by the S = 2048 / 2049
// comparison) into the epilogue -- its score as a 64-long dot product on the CUDA cores, one more term in the merge of the key
// halves.  Parity-green, 32 steps instead of 33, and slower in every placement tried: 148 -> 152-153 us at batch 8 with the
// dot product in the epilogue (operands staged in shared memory) or before the loop (loads overlapping the pipeline fill).)
// (Tried and archived as tools/attention_persist_experiment.cu.txt: a persistent grid over a work counter, producer and
// issuer running ahead across tile boundaries so that the MMA pipeline never drains.  Parity-green, not faster: 168.6 us at
// batch 8, 612 us at batch 32 -- the co-resident CTA already fills a CTA's prologue and epilogue.)
// (Tried and removed: pacing the issuing thread between MMAs.  An 80-cycle clock spin after each MMA removes the lag of the
// issuer's scheduler mates completely -- all eight softmax warps then arrive within 150 cycles of each other -- but the
// issuer itself becomes the limit (950 cycles per step, 176 us); 20-55 cycle pauses change nothing, 146.8 us.)
__device__ unsigned int g_att_sm_slots[1024];      // per SM: bit s set while a resident CTA holds slot s
constexpr uint32_t ATS_COL_S = 0, ATS_COL_O = 128;
constexpr int ATS_SMEM = ATT_SMEM + 2 * 128 * 8;

template <int POLY>
__global__ void __launch_bounds__(ATS_THREADS, 2)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse_out, const float* __restrict__ kmax2, int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = sK + K_STAGES * TILE_BYTES;
  AttBarriers* bars = reinterpret_cast<AttBarriers*>(sV + V_STAGES * TILE_BYTES);
  float2* ml = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 halves][128 rows] (m, l)

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = b * S;
  const int ntiles = (S + KT - 1) / KT;
  const int nsub = (S + KS - 1) / KS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < K_STAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 2); }
    for (int s = 0; s < V_STAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 8);
      mbar_init(&bars->pv_done[i], 1);
    }
    fence_mbar_init();
  }
  __shared__ int s_slot;
  unsigned int smid = 0;
  if (threadIdx.x == 0) {
    int slot = 0;
    if (kSingleIssuer && HSENET_ATT_SPREAD) {
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      smid &= 1023u;
      if (atomicOr(&g_att_sm_slots[smid], 1u) & 1u) {      // slot 0 taken by the co-resident CTA
        atomicOr(&g_att_sm_slots[smid], 2u);
        slot = 1;
      }
    }
    s_slot = slot;
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
  const int slot = __shfl_sync(0xffffffffu, s_slot, 0);
  const int producer_warp = 2 * slot, issuer_warp = 2 * slot + 1;
  if (threadIdx.x == 0) { ATT_TV(11, 47, issuer_warp); }
  pdl_prologue_done();

  auto mma_issuer = [&](const int par) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(QT, KS, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(QT, kHeadDim, 0, 1);
    const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
    auto issue_qk = [&](const int t) {
      const int ks = (t >> 1) % K_STAGES, sub = t & 1;
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + ks * TILE_BYTES + sub * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + ATS_COL_S + sub * KS, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        tc_commit(&bars->s_full[sub]);
        tc_commit(&bars->k_empty[ks]);
      }
      __syncwarp();
    };
    if (kSingleIssuer) {
      // ONE issuing warp for every step: P V (t) then Q K^T (t+2), all MMAs of the CTA in program order.  With two issuing
      // warps (round 1: needed because a diverged single-lane issuer cost ~1200 cycles per step) P V (t) had to wait for
      // the COMPLETION of the other warp's P V (t-1) to keep the accumulation order fixed, which put a full MMA round trip
      // (tools/attn_timeline.py: 500-950 cycles from the softmax warps' arrive to the issuer's wake-up) on the cycle.
      if (par != 0) return;
      ctl_wait(&bars->q_full, 0);
      ctl_wait(&bars->k_full[0], 0);
      issue_qk(0);
      if (nsub > 1) issue_qk(1);
      ATT_TR_DECL;
      for (int t = 0; t < nsub; ++t) {
        const int sub = t & 1, j = t >> 1, vs = j % V_STAGES, t2 = t + 2;
        if (lane == 0) { ATT_TS(12, t); }
        ctl_wait(&bars->v_full[vs], (j / V_STAGES) & 1);
        if (lane == 0) { ATT_TS(13, t); }
        if (t2 < nsub) ctl_wait(&bars->k_full[(t2 >> 1) % K_STAGES], ((t2 >> 1) / K_STAGES) & 1);
        ATT_TR(0);
        if (lane == 0) { ATT_TS(11, t); }
        ctl_wait(&bars->p_full[sub], (t >> 1) & 1);
        if (lane == 0) { ATT_TS(9, t); }
        tc_fence_after();
        ATT_TR(1);
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES + sub * SUB_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < KS / 16; ++i) {
            const int k = kPvInterleave ? ((i & 1) * 2 + (i >> 1)) : i;
            const int half = k >> 1;
            umma_ts(tmem_base + ATS_COL_O + half * kHeadDim, tmem_base + ATS_COL_S + sub * KS + half * 32 + 8 * (k & 1),
                    vdesc + 128 * k, idesc_pv, (t | (k & 1)) != 0 ? 1u : 0u);
          }
          tc_commit(&bars->pv_done[sub]);
          tc_commit(&bars->v_empty[vs]);
        }
        __syncwarp();
        ATT_TR(2);
        if (t2 < nsub) issue_qk(t2);
        if (lane == 0) { ATT_TS(10, t); }
        ATT_TR(3);
      }
      ATT_TR_DUMP(16);
      return;
    }
    if (par >= nsub) return;
    ctl_wait(&bars->q_full, 0);
    ctl_wait(&bars->k_full[0], 0);
    issue_qk(par);
    ATT_TR_DECL;
    for (int t = par; t < nsub; t += 2) {
      const int j = t >> 1, vs = j % V_STAGES, t2 = t + 2;
      ctl_wait(&bars->v_full[vs], (j / V_STAGES) & 1);
      if (t2 < nsub) ctl_wait(&bars->k_full[(t2 >> 1) % K_STAGES], ((t2 >> 1) / K_STAGES) & 1);
      if (t >= 1) ctl_wait(&bars->pv_done[par ^ 1], ((t - 1) >> 1) & 1);   // keep the accumulation order fixed
      ATT_TR(0);
      ctl_wait(&bars->p_full[par], (t >> 1) & 1);
      if (lane == 0) { ATT_TS(9, t); }
      tc_fence_after();
      ATT_TR(1);
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES + par * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < KS / 16; ++i) {
          // key half k>>1: A = 16 keys = 8 packed columns at the start of that half's score columns, D = that half's O.
          // Issue order k = 0, 2, 1, 3: consecutive MMAs alternate between the two accumulators (back-to-back MMAs into
          // the SAME accumulator occupy the pipe ~57 cycles instead of ~48, tools/umma_throughput.cu)
          const int k = kPvInterleave ? ((i & 1) * 2 + (i >> 1)) : i;
          const int half = k >> 1;
          umma_ts(tmem_base + ATS_COL_O + half * kHeadDim, tmem_base + ATS_COL_S + par * KS + half * 32 + 8 * (k & 1),
                  vdesc + 128 * k, idesc_pv, (t | (k & 1)) != 0 ? 1u : 0u);
        }
        tc_commit(&bars->pv_done[par]);
        tc_commit(&bars->v_empty[vs]);
      }
      __syncwarp();
      ATT_TR(2);
      if (t2 < nsub) issue_qk(t2);
      if (lane == 0) { ATT_TS(10, t); }
      ATT_TR(3);
    }
    ATT_TR_DUMP(16 + 16 * par);
  };

  if (warp == producer_warp) {
    // TMA producer (whole warp, elect-guarded issue; K one tile ahead of V -- see attention_kernel)
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, &bars->q_full, h * kHeadDim, row0 + q0);
    }
    __syncwarp();
    auto load_k = [&](const int j) {
      const int ks = j % K_STAGES;
      ctl_wait(&bars->k_empty[ks], ((j / K_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->k_full[ks], TILE_BYTES);
        tma_load_2d_hint(sK + ks * TILE_BYTES, &tmQKV, &bars->k_full[ks], kHidden + h * kHeadDim, row0 + j * KT,
                         kEvictLast);
      }
      __syncwarp();
    };
    load_k(0);
    for (int j = 0; j < ntiles; ++j) {
      const int vs = j % V_STAGES;
      if (j + 1 < ntiles) load_k(j + 1);
      ctl_wait(&bars->v_empty[vs], ((j / V_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->v_full[vs], TILE_BYTES);
        tma_load_2d_hint(sV + vs * TILE_BYTES, &tmQKV, &bars->v_full[vs], 2 * kHidden + h * kHeadDim,
                         row0 + j * KT, kEvictLast);
      }
      __syncwarp();
      // This warp has slack: it also observes the P V completions of the previous tile's two steps, so that every phase of
      // pv_done has a waiter before the barrier is committed again (compute-sanitizer synccheck's rule; the softmax warps
      // only wait on it in the rare rescale path and at the end).  On the issuing warp the same wait cost 5 %.  The V stage
      // this loop waits for next is released by the same two P V groups, so nothing is delayed.
      if (kSingleIssuer && HSENET_ATT_OBSERVE_PV && j >= 1) {
        ctl_wait(&bars->pv_done[0], (j - 1) & 1);
        if (2 * (j - 1) + 1 < nsub) ctl_wait(&bars->pv_done[1], (j - 1) & 1);
      }
    }
  } else if (warp == issuer_warp) {
    mma_issuer(0);
  } else if (!kSingleIssuer && warp == 2) {
    mma_issuer(1);        // two-issuer experiment (slot is always 0 there: producer warp 0, issuers 1 and 2)
  } else if (warp >= 4) {
    // ===================== softmax warps: quarter = TMEM lane quarter (warp % 4), half = key half =====================
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int qi = q0 + quarter * 32 + lane;
    const bool warp_live = (q0 + quarter * 32) < S;
    const float c = 0.125f * 1.4426950408889634f;
    const uint32_t col_o = ATS_COL_O + half * kHeadDim;
    float m = -INFINITY;
    float l = 0.f;
    // MAX-FREE mode.  By Cauchy-Schwarz every scaled score of this row is <= mhat = c |q| max_k |k|, so mhat can replace the
    // running maximum: P = 2^(s c - mhat) never overflows, the row sum / output ratio is unchanged, and the per-step row
    // maximum (0.8 of ~4.5 instructions per score), the lazy-rescale test, the warp vote and the O correction all vanish.
    // bf16 / fp32 carry an 8-bit exponent, so a loose bound costs no precision as long as nothing underflows: the worst
    // case is a score of -mhat (q anti-aligned with a key), i.e. P = 2^(-2 mhat), hence the mode is only taken when
    // mhat <= 50 for every row of the warp (true for LayerNorm'ed activations; otherwise the exact online softmax runs).
    bool bounded = false;
    if (kmax2 != nullptr) {
      smx_wait(&bars->q_full, 0);
      const uint4* qrow = reinterpret_cast<const uint4*>(sQ + (quarter * 32 + lane) * 128);   // swizzle permutes chunks only
      float q2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 u = qrow[j];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
          q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
        }
      }
      const float mhat = c * sqrtf(q2 * __ldg(kmax2 + b * kHeads + h)) * 1.0001f;
      bounded = !__any_sync(0xffffffffu, !(mhat <= 50.0f));
      if (bounded) m = mhat;
    }
    ATT_TR_DECL;
    auto softmax_step = [&](const int t, auto masked, auto maxfree) {
      constexpr bool MAXFREE = decltype(maxfree)::value;
      const int bsel = t & 1;
      const uint32_t col_s = ATS_COL_S + bsel * KS + half * 32;
      smx_wait(&bars->s_full[bsel], (t >> 1) & 1);
      ATT_TR(0);
      if (warp == 4 && lane == 0) { ATT_TS(8, t); }
      tc_fence_after();
      ATT_TR(2);
      uint32_t x[32];
      uint32_t pk[16];
      float alpha = 1.f;
      if (warp_live) tmem_ld32(tmem_base + lane_base + col_s, x);
      if (warp_live) {
        tmem_ld_wait();
        ATT_TR(1);
        if constexpr (decltype(masked)::value) {
          const int kbase = t * KS + half * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + i >= S) x[i] = 0xff800000u;
        }
        if constexpr (!MAXFREE) {
          float mx[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = __uint_as_float(x[u]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(x[i + u]));
          }
          const float tm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c;
          if (tm > m + 8.0f) {
            alpha = ex2(m - tm);
            m = tm;
          }
        }
        ATT_TR(3);
        // a half that has not seen a valid key yet (only possible in a masked step) keeps m = -inf: subtract 0 instead,
        // so that the all -inf scores give exp2(-inf) = 0 and not exp2(-inf + inf)
        float msub = m;
        if constexpr (decltype(masked)::value && !MAXFREE) msub = (m == -INFINITY) ? 0.f : m;
        const uint64_t c2 = pack2(c, c), nm2 = pack2(-msub, -msub);
        uint64_t rs2 = pack2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float e[8];
#pragma unroll
          for (int u = 0; u < 8; u += 2)
            unpack2(fma2(pack2(__uint_as_float(x[i + u]), __uint_as_float(x[i + u + 1])), c2, nm2), e[u], e[u + 1]);
#pragma unroll
          for (int u = 0; u < 8 - POLY; ++u) e[u] = ex2(e[u]);
#pragma unroll
          for (int u = 8 - POLY; u < 8; u += 2) exp2_poly2(e[u], e[u + 1], e[u], e[u + 1]);
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            rs2 = add2(rs2, pack2(e[u], e[u + 1]));
            pk[(i + u) >> 1] = pack_bf16x2(e[u], e[u + 1]);
          }
        }
        float rs0, rs1;
        unpack2(rs2, rs0, rs1);
        if constexpr (MAXFREE) l += rs0 + rs1; else l = l * alpha + (rs0 + rs1);
        ATT_TR(4);
        tmem_st16(tmem_base + lane_base + col_s, pk);
        ATT_TR(5);
      }
      if (!MAXFREE && t > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        smx_wait(&bars->pv_done[bsel ^ 1], ((t - 1) >> 1) & 1);
        if (t >= 2) smx_wait(&bars->pv_done[bsel], ((t >> 1) - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          tmem_ld32(tmem_base + lane_base + col_o + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_base + lane_base + col_o + ch * 32, o);
        }
      }
      ATT_TR(6);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[bsel]);
      if (lane == 0) { ATT_TS(warp - 4, t); }
      ATT_TR(7);
    };
    const bool ragged = (S % KS) != 0;
    if (bounded) {
      for (int t = 0; t < nsub - (ragged ? 1 : 0); ++t) softmax_step(t, std::false_type{}, std::true_type{});
      if (ragged) softmax_step(nsub - 1, std::true_type{}, std::true_type{});
    } else {
      for (int t = 0; t < nsub - (ragged ? 1 : 0); ++t) softmax_step(t, std::false_type{}, std::false_type{});
      if (ragged) softmax_step(nsub - 1, std::true_type{}, std::false_type{});
    }
    if (warp == 4 && lane == 0) { ATT_TR_DUMP(0); }
    // ---- merge the two key halves and store: this warp takes output columns [half*32, half*32+32) of its 32 rows ----
    ml[half * 128 + quarter * 32 + lane] = make_float2(m, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");                  // the 8 softmax warps only
    const float2 other = ml[(half ^ 1) * 128 + quarter * 32 + lane];
    const float ms = fmaxf(m, other.x);
    const float w_self = ex2(m - ms), w_other = ex2(other.x - ms);
    const float lsum = l * w_self + other.y * w_other;
    const float inv = 1.0f / lsum;
    if (lse_out != nullptr && half == 0) {
      const int sp = ((S + QT - 1) / QT) * QT;
      lse_out[(static_cast<long>(b) * kHeads + h) * sp + qi] = qi < S ? ms + log2f(lsum) : INFINITY;
    }
    const float fa = (half == 0 ? w_self : w_other) * inv;          // weight of O_A
    const float fb = (half == 0 ? w_other : w_self) * inv;          // weight of O_B
    if (nsub >= 2) smx_wait(&bars->pv_done[(nsub - 2) & 1], ((nsub - 2) >> 1) & 1);
    smx_wait(&bars->pv_done[(nsub - 1) & 1], ((nsub - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t oa[32], ob[32];
    tmem_ld32(tmem_base + lane_base + ATS_COL_O + half * 32, oa);
    tmem_ld32(tmem_base + lane_base + ATS_COL_O + kHeadDim + half * 32, ob);
    tmem_ld_wait();
    if (qi < S) {
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long>(row0) + qi) * kHidden + h * kHeadDim + half * 32);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = __uint_as_float(oa[g * 8 + u]) * fa + __uint_as_float(ob[g * 8 + u]) * fb;
        uint4 u4;
        u4.x = pack_bf16x2(v[0], v[1]);
        u4.y = pack_bf16x2(v[2], v[3]);
        u4.z = pack_bf16x2(v[4], v[5]);
        u4.w = pack_bf16x2(v[6], v[7]);
        dst[g] = u4;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0 && kSingleIssuer && HSENET_ATT_SPREAD) atomicAnd(&g_att_sm_slots[smid], ~(1u << slot));
}

}  // namespace

int attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, float* kmax_scratch, int B, int S,
                   cudaStream_t stream) {
  if (B <= 0 || S <= 0) return HS_OK;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return HS_ERR_ALIGN;
  CUtensorMap tm;
  const int rc = make_tmap_2d_bf16(&tm, qkv, 3 * kHidden, static_cast<uint64_t>(B) * S, 3 * kHidden, kHeadDim, 128);
  if (rc != HS_OK) return rc;
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    bool ok = true;
    ok &= cudaFuncSetAttribute(attention_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_split_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_split_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attention_split_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM) == cudaSuccess;
    if (!ok) return HS_ERR_CUDA;
  }
  ProfScope prof(PROF_ATTENTION, 4.0 * B * kHeads * double(S) * S * kHeadDim, 2.0 * B * double(S) * 4 * kHidden,
                 stream);
  const char* e = std::getenv("HSENET_ATT_POLY");        // share of exponentials emulated on the FMA pipe (0 / 2 / 4 of 8)
  const int poly = e != nullptr ? std::atoi(e) : kDefaultPoly;
  // HSENET_ATT_KERNEL (A/B runs): "split" (default) = 8 softmax warps splitting every step by key half (two O accumulators);
  // "tri" = 4 softmax warps, three aliased S/P buffers; "rowwarp" = the round-1 layout (two S + two P buffers)
  const char* kv = std::getenv("HSENET_ATT_KERNEL");
  const bool split = kv == nullptr || kv[0] == 's';
  const bool tri = kv != nullptr && kv[0] == 't';
  const dim3 grid((S + QT - 1) / QT, kHeads, B);
  if (split) {
    // Max-free softmax (OPT-IN, HSENET_ATT_MAXFREE=1): needs the largest key norm per (volume, head), a small pre-pass over K
    // into the caller's scratch.  Measured on B200: it removes 26 % of the softmax warps' instructions (row max, rescale
    // test, vote, O correction) and the kernel is NOT faster (148.0 -> 155.8 us at batch 8 with the pre-pass, equal without
    // it) -- the step time is not set by the softmax instruction stream (profiles/README.md) -- so the exact online softmax
    // stays the default.
    const char* mf = std::getenv("HSENET_ATT_MAXFREE");
    const float* kmax = nullptr;
    if (kmax_scratch != nullptr && mf != nullptr && mf[0] == '1') {
      if (cudaMemsetAsync(kmax_scratch, 0, static_cast<size_t>(B) * kHeads * sizeof(float), stream) != cudaSuccess)
        return HS_ERR_CUDA;
      attn_kmax_kernel<<<dim3((S + 15) / 16, B), dim3(kHeads, 16), 0, stream>>>(qkv, reinterpret_cast<int*>(kmax_scratch), S);
      count_launch();
      kmax = kmax_scratch;
    }
    if (poly >= 4) launch_pdl(attention_split_kernel<4>, grid, dim3(ATS_THREADS), ATS_SMEM, stream, tm, out, lse, kmax, S);
    else if (poly >= 2) launch_pdl(attention_split_kernel<2>, grid, dim3(ATS_THREADS), ATS_SMEM, stream, tm, out, lse, kmax, S);
    else launch_pdl(attention_split_kernel<0>, grid, dim3(ATS_THREADS), ATS_SMEM, stream, tm, out, lse, kmax, S);
  } else if (tri) {
    if (poly >= 4) launch_pdl(attention_kernel<4, 3>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
    else if (poly >= 2) launch_pdl(attention_kernel<2, 3>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
    else launch_pdl(attention_kernel<0, 3>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
  } else {
    if (poly >= 4) launch_pdl(attention_kernel<4, 2>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
    else if (poly >= 2) launch_pdl(attention_kernel<2, 2>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
    else launch_pdl(attention_kernel<0, 2>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, out, lse, S);
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs

#ifdef HSENET_ATT_TRACE
extern "C" int hsenet_debug_att_times(long long* host672) {
  return cudaMemcpyFromSymbol(host672, hs::g_att_times, sizeof(long long) * 672) == cudaSuccess ? 0 : -4;
}
extern "C" int hsenet_debug_att_trace(unsigned long long* host48) {
  return cudaMemcpyFromSymbol(host48, hs::g_att_trace, sizeof(unsigned long long) * 48) == cudaSuccess ? 0 : -4;
}
#endif
