// conv3d_igemm.cu — stride-1 3-D convolution forward and data-gradient as an implicit GEMM on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA.
//
//   D[m][n] = sum_k A[m][k] * B[n][k]      m = output voxel (n,t,h,w), n = output channel,
//                                           k = (segment, tap, input channel)
//
// A is never materialised: one k-block = 64 channels of ONE filter tap, and the 128 voxels of an M tile
// form a (bw x bh x bt) box of the NDHWC activation tensor, so the A tile of tap (it,ih,iw) is the SAME
// box shifted by (it-pt, ih-ph, iw-pw) — a single 5-D TMA tile load whose out-of-bounds elements are
// zero-filled by the hardware. That zero fill *is* the convolution padding, including the causal
// (front-only) time padding of CausalConv3d: pt = kt-1 simply shifts the box start to negative t.
// The reference does this with an explicit F.pad copy followed by conv3d (genie/module/video.py:160-192).
//
// Forward uses K-major weights w[cout][k]; the data gradient reuses the SAME weight buffer as an
// MN-major B operand (k = cout rows of 64, n = cin contiguous) with mirrored taps, so no transposed
// weight copy ever exists.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0    TMA producer (one elected lane)
//   warp 1    TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5 epilogue: TMEM -> registers -> (+bias) -> global, double-buffered against the next tile's MMAs
#include <stdlib.h>

#include <utility>
#include <vector>

#include "og_host.cuh"
#include "og_ptx.cuh"

namespace og {
extern std::atomic<uint64_t> g_launches;

// One K segment = a (sub)set of filter taps over one input tensor. Along every axis d the taps visited are
// tap0[d] + j*tstep[d] (j < n[d]) of a (.., kh, kw) filter, and the A box of tap j is the M-tile box shifted by
// sh0[d] + j*shstep[d]:
//   forward                 n = k, tap0 = 0, tstep = 1, sh0 = -pad, shstep = +1      (x[v*stride + tap - pad])
//   data gradient, stride 1 n = k, tap0 = 0, tstep = 1, sh0 = +pad, shstep = -1      (dy[v - (tap - pad)])
//   data gradient, stride s one launch per residue class of the input position: n = #taps of that class,
//                           tap0 = class, tstep = s, sh0 = e, shstep = -1            (see og_conv3d_strided_dgrad)
// axis order: 0 = t, 1 = h, 2 = w.
struct IgemmSeg {
  int cin_blocks;  // channels / 64
  int n[3], tap0[3], tstep[3], sh0[3], shstep[3];
  int kh, kw;      // filter extents (tap index = (it*kh + ih)*kw + iw)
};

struct IgemmParams {
  int nseg;
  IgemmSeg seg[2];
  int sx[3];       // A box start = M-tile start * sx + shift (forward strided convolution: sx = stride; else 1)
  int OT, OH, OW;  // extents of the OUTPUT tensor; M-tile voxel (t,h,w) is stored at (t*om[0]+oo[0], h*om[1]+oo[1], ...)
  int om[3], oo[3];
  int b_mn_major;  // 0: B tile is [n rows][64 k] (K-major); 1: B tile is [k rows][n] in 64-wide panels (MN-major)
  int block_n;     // UMMA N (16..256)
  int num_n_tiles, num_m_tiles;
  int bw_log2, bh_log2, bt_log2, bn_log2;  // M tile = 2^bw x 2^bh x 2^bt x 2^bn voxels (product 128)
  int tiles_w, tiles_h, tiles_t;           // M tiles along each axis (ceil)
  int N, T, H, W;
  int n_out;      // valid output channels
  long long ldo;  // output row stride in elements
  void* out;
  int out_f32;
  int vec_ok;  // rows are 16-byte aligned and n_out is a multiple of the vector width
  const float* bias0;
  const float* bias1;
  int num_stages;
  int num_kb;  // k-blocks per tile
  int m_sub;   // M sub-tiles (of 128 rows) per CTA tile sharing one B stage: 1 or 2
  int fast_store;  // bf16 output, n_out % 64 == 0, block_n % 64 == 0: coalesced staged stores
  const __nv_bfloat16* residual;  // optional bf16 [voxels][n_out] added to the output (out = conv + bias + residual)
  // fused epilogue reductions (fast_store path, one sample per CTA tile):
  double* gn_sums;             // forward: [N][2] += (sum y, sum y^2) of the bf16-rounded output (GroupNorm(1,C) statistics)
  const __nv_bfloat16* red_x;  // data gradient: GN input x [voxels][n_out] ...
  const float* red_A;          // ... per-(sample, channel) scale / shift of y = act(x*A+B) ...
  const float* red_B;
  float* red_S;                // ... S[n][c] += (sum dpre, sum dpre*x), dpre = d * act'(x*A+B)   (og_affine_act_bwd_reduce)
  int red_act;
  int swap;        // operand swap (see below): D^T[cout][256 voxels] = W . X^T, used when Cout tiles are 128 wide
  unsigned int* sched;  // dynamic tile scheduler state {magic, next item, CTAs done} in the caller's workspace, or NULL
  int splits;      // split-K factor (1 = none): each work item covers a k-block range and reduces into `ws`
  float* ws;       // fp32 [voxels][n_out] partial-sum workspace (zero on entry) when splits > 1
  long long ws_slab;  // > 0: every split item STORES its partial tile into its own slab ws[split*ws_slab + ...] (no atomics,
                      // no memset; the finish pass adds the slabs); 0: red.global.add into one zeroed slab
  int dbg;         // timing experiments only (OG_IGEMM_DBG): 1 = skip the split-K epilogue's global writes
  unsigned int* tile_ctr;  // per-tile arrival counters (prepared workspace): the LAST split item of a tile finishes it in-kernel
                           // (bias, cast, store, GroupNorm sums) and re-zeroes its part of `ws` — no memset, no finish launch
};

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;                       // 64 bf16 = one 128-byte swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KiB
static constexpr int kTmemCols = 512;
static constexpr int kAccStride = 256;                   // columns between the two accumulator buffers
static constexpr int kMaxStages = 8;
static constexpr int kThreads = 192;
static constexpr int kSchedDepth = 4;
static constexpr unsigned int kSchedMagic = 0x0695CED0u;   // written by og_workspace_init

struct TileCoord {
  int n0, t0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_m_tile(const IgemmParams& p, int m_tile) {
  TileCoord c;
  int per_sample = p.tiles_w * p.tiles_h * p.tiles_t;
  const int tn = m_tile / per_sample;
  c.n0 = tn << p.bn_log2;
  int r = m_tile - tn * per_sample;
  int tw = r % p.tiles_w;
  r /= p.tiles_w;
  int th = r % p.tiles_h;
  int tt = r / p.tiles_h;
  c.w0 = tw << p.bw_log2;
  c.h0 = th << p.bh_log2;
  c.t0 = tt << p.bt_log2;
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
    og_conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                         const __grid_constant__ CUtensorMap mapB, const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128-byte swizzle pattern shared by TMA and the UMMA descriptors
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.block_n * kBlockK * 2;
  const int a_bytes = p.m_sub * kABytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.num_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  // Dynamic tile scheduler (round 2). With the static `item = blockIdx.x + i * gridDim.x` assignment a CTA that cannot be
  // resident (another kernel — the NCCL all-reduce of the data-parallel step — holds its SM; this kernel's 227 KB of
  // shared memory exclude co-residency) starts only when a sibling CTA finishes and then still owns its full share of
  // tiles: the launch takes ~2x as long. Here the producer warp draws work items from a global counter and hands them to
  // the MMA / epilogue warps through a 4-deep shared-memory ring, so late CTAs simply find no work left.
  uint64_t* sched_full = bars + 2 * kMaxStages + 5;    // [4]
  uint64_t* sched_empty = bars + 2 * kMaxStages + 9;   // [4]
  int* sched_item = reinterpret_cast<int*>(bars + 2 * kMaxStages + 13);  // [4]
  const bool dyn = p.sched != nullptr && p.sched[0] == kSchedMagic;
  float* bias_s = reinterpret_cast<float*>(bars + 32);                    // [256] bias0+bias1 of the current N tile
  uint8_t* stage_s = reinterpret_cast<uint8_t*>(bars + 32) + 1024;        // 4 warps x 32 rows x 128 B store staging
  float* coef_s = reinterpret_cast<float*>(stage_s + 4 * 4096);          // [2][256] A, B of the current sample / N tile
  float* red_s = coef_s + 512;                                            // [256][2] column sums of the current tile
  double* stat_s = reinterpret_cast<double*>(red_s + 512);                // [2]
  int* flag_s = reinterpret_cast<int*>(stat_s + 2);                       // [1] "this CTA finishes the tile" (fused split-K)

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int num_m_super = (p.num_m_tiles + p.m_sub - 1) / p.m_sub;
  const int total_tiles = num_m_super * p.num_n_tiles * p.splits;  // work items

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nseg > 1) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    for (int a = 0; a < kSchedDepth; ++a) {
      mbar_init(&sched_full[a], 1);
      mbar_init(&sched_empty[a], 5);   // MMA warp + 4 epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (all 32 lanes run the loop converged; one elected lane issues — see elect_one() in og_ptx.cuh)
    {
      int stage = 0;
      uint32_t phase = 0;
      int sslot = 0;
      uint32_t sphase = 0;
      for (int iter = 0;; ++iter) {
        int item;
        if (dyn) {
          mbar_wait(&sched_empty[sslot], sphase ^ 1);
          int got = 0;
          if (elect_one()) got = (int)atomicAdd(p.sched + 1, 1u);
          item = __reduce_max_sync(0xffffffffu, got);          // broadcast (items are >= 0)
          if (elect_one()) {
            sched_item[sslot] = item;
            mbar_arrive(&sched_full[sslot]);                    // release: the slot is visible to the waiters
          }
          __syncwarp();
          if (++sslot == kSchedDepth) {
            sslot = 0;
            sphase ^= 1;
          }
        } else {
          item = blockIdx.x + iter * gridDim.x;
        }
        if (item >= total_tiles) break;
        const int tile = item / p.splits, split = item - tile * p.splits;
        const int m_super = tile / p.num_n_tiles;
        const int n_tile = tile - m_super * p.num_n_tiles;
        const TileCoord tc = decode_m_tile(p, m_super * p.m_sub);
        const TileCoord tc1 = decode_m_tile(p, m_super * p.m_sub + 1);  // second sub-tile (m_sub == 2)
        const int kb_begin = (int)(((long long)p.num_kb * split) / p.splits);
        const int kb_end = (int)(((long long)p.num_kb * (split + 1)) / p.splits);
        // position (segment, tap = (jt, jh, jw), channel block) of kb_begin: divisions once per work item,
        // then the single producer thread only increments with carries (its instruction count per k-block
        // is what bounds the pipeline when MMAs are short)
        const int nkb0 = p.seg[0].cin_blocks * p.seg[0].n[0] * p.seg[0].n[1] * p.seg[0].n[2];
        int sidx = (kb_begin >= nkb0 && p.nseg > 1) ? 1 : 0;
        IgemmSeg sg = p.seg[sidx];
        int rel = kb_begin - (sidx ? nkb0 : 0);
        int tapc = rel / sg.cin_blocks;
        int cb = rel - tapc * sg.cin_blocks;
        int jt = tapc / (sg.n[1] * sg.n[2]);
        int jh = (tapc / sg.n[2]) % sg.n[1];
        int jw = tapc % sg.n[2];
        const int aw0 = tc.w0 * p.sx[2], ah0 = tc.h0 * p.sx[1], at0 = tc.t0 * p.sx[0];
        const int bw0 = tc1.w0 * p.sx[2], bh0 = tc1.h0 * p.sx[1], bt0 = tc1.t0 * p.sx[0];
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const CUtensorMap* mapA = (sidx == 0) ? &mapA0 : &mapA1;
          const int dt = sg.sh0[0] + jt * sg.shstep[0], dh = sg.sh0[1] + jh * sg.shstep[1],
                    dw = sg.sh0[2] + jw * sg.shstep[2];
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (elect_one()) {
            mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
            tma_load_5d(sa, mapA, &full[stage], cb * kBlockK, aw0 + dw, ah0 + dh, at0 + dt, tc.n0);
            if (p.m_sub == 2)
              tma_load_5d(sa + kABytes, mapA, &full[stage], cb * kBlockK, bw0 + dw, bh0 + dh, bt0 + dt, tc1.n0);
            if (!p.b_mn_major) {
              tma_load_2d(sb, &mapB, &full[stage], kb * kBlockK, n_tile * p.block_n);
            } else {
              // w[co][tap][ci] as (ci, tap, co): one (64 ci, 1 tap, 64 co) box per 64-wide N panel
              const int tap = ((sg.tap0[0] + jt * sg.tstep[0]) * sg.kh + sg.tap0[1] + jh * sg.tstep[1]) * sg.kw +
                              sg.tap0[2] + jw * sg.tstep[2];
              for (int pp = 0; pp < p.block_n / 64; ++pp)
                tma_load_3d(sb + pp * (64 * 128), &mapB, &full[stage], n_tile * p.block_n + pp * 64, tap,
                            cb * kBlockK);
            }
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++cb == sg.cin_blocks) {
            cb = 0;
            if (++jw == sg.n[2]) {
              jw = 0;
              if (++jh == sg.n[1]) {
                jh = 0;
                if (++jt == sg.n[0]) {  // next segment (the fused 1x1x1 shortcut)
                  sidx = 1;
                  sg = p.seg[1];
                  jt = jh = jw = 0;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
    if (dyn && elect_one()) {
      // last CTA to finish resets the counter for the next launch that uses this workspace (stream-ordered)
      __threadfence();
      const unsigned int old = atomicInc(p.sched + 2, gridDim.x - 1);
      if (old == gridDim.x - 1) {
        __threadfence();
        p.sched[1] = 0u;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, (uint32_t)p.block_n, 0u, (uint32_t)p.b_mn_major);
      const uint32_t idesc_swap = umma_idesc_bf16(kBlockM, 256u, (uint32_t)p.b_mn_major, 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int sslot = 0;
      uint32_t sphase = 0;
      for (int iter = 0;; ++iter) {
        int item;
        if (dyn) {
          mbar_wait(&sched_full[sslot], sphase);
          item = sched_item[sslot];
          __syncwarp();
          if (elect_one()) mbar_arrive(&sched_empty[sslot]);
          if (++sslot == kSchedDepth) {
            sslot = 0;
            sphase ^= 1;
          }
        } else {
          item = blockIdx.x + iter * gridDim.x;
        }
        if (item >= total_tiles) break;
        const int split = item % p.splits;
        const int nkb = (int)(((long long)p.num_kb * (split + 1)) / p.splits) -
                        (int)(((long long)p.num_kb * split) / p.splits);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
          if (p.swap) {
            // Swapped operands: A = the 128 x 64 weight tile (M = output channels), B = the TWO voxel sub-tiles as one
            // 256-row K-major operand (they sit back to back in the stage). One 128x256x16 MMA replaces two
            // 128x128x16 ones: the same bytes per stage, but 96 instead of 128 B/clk of shared-memory operand reads,
            // which is what capped the Cout = 128 layers at ~78 % of the Cout = 256 rate.
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t wdesc = p.b_mn_major ? umma_smem_desc_sw128(b_addr + k * 2048, 64 * 128, 1024)
                                                    : umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                const uint64_t xdesc = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                umma_bf16_ss(d_tmem, wdesc, xdesc, idesc_swap, (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty[stage]);
            }
            __syncwarp();
          } else if (elect_one()) {
            for (int ms = 0; ms < p.m_sub; ++ms) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // A: K-major, 128-byte rows, 8-row groups 1024 B apart; advance 16 elements = 32 B inside the row
                const uint64_t adesc = umma_smem_desc_sw128(a_addr + ms * kABytes + k * 32, 16, 1024);
                uint64_t bdesc;
                if (!p.b_mn_major) {
                  bdesc = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                } else {
                  // B: MN-major panels [block_n/64][64 k-rows][128 B]; 16 k-rows = 2048 B; panel stride 8192 B
                  bdesc = umma_smem_desc_sw128(b_addr + k * 2048, 64 * 128, 1024);
                }
                umma_bf16_ss(d_tmem + ms * p.block_n, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          }
          if (!p.swap) __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float fl_s = 0.f, fl_ss = 0.f;
    for (int j = threadIdx.x - 64; j < 512; j += 128) red_s[j] = 0.f;
    if (threadIdx.x == 64) stat_s[0] = stat_s[1] = 0.0;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    int sslot = 0;
    uint32_t sphase = 0;
    for (int iter = 0;; ++iter) {
      int item;
      if (dyn) {
        mbar_wait_relaxed(&sched_full[sslot], sphase);
        item = sched_item[sslot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sched_empty[sslot]);
        if (++sslot == kSchedDepth) {
          sslot = 0;
          sphase ^= 1;
        }
      } else {
        item = blockIdx.x + iter * gridDim.x;
      }
      if (item >= total_tiles) break;
      const int tile = item / p.splits;
      const int m_super = tile / p.num_n_tiles;
      const int n_tile = tile - m_super * p.num_n_tiles;
      mbar_wait_relaxed(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.swap) {
        // accumulator is D^T: TMEM lane = output channel (this thread owns channel `co`), column = voxel of the
        // 256-voxel super tile. 32-voxel chunks are transposed through a [32 voxels][128 channels] bf16 staging tile
        // shared by the four epilogue warps (double-buffered, one named barrier per chunk), then stored as full
        // 256-byte voxel rows. Host guarantees: whole boxes, n_out % 128 == 0, no residual / split-K.
        const TileCoord tcs0 = decode_m_tile(p, m_super * 2), tcs1 = decode_m_tile(p, m_super * 2 + 1);
        const int co = n_tile * 128 + row;
        float bsum = 0.f;
        if (p.bias0) bsum += __ldg(p.bias0 + co);
        if (p.bias1) bsum += __ldg(p.bias1 + co);
        const uint32_t t_addr = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
        float st_s = 0.f, st_ss = 0.f;
        for (int c = 0; c < 256; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c, v);
          tmem_ld_wait();
          uint8_t* buf = stage_s + ((c >> 5) & 1) * 8192;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(__uint_as_float(v[j]) + bsum);
            if (p.gn_sums) {
              const float r = __bfloat162float(hb);
              st_s += r;
              st_ss = fmaf(r, r, st_ss);
            }
            *reinterpret_cast<__nv_bfloat16*>(buf + j * 256 + row * 2) = hb;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = (warp - 2) * 8 + i * 2 + (lane >> 4);
            const int vv = c + j;
            const TileCoord tc = (vv >> 7) ? tcs1 : tcs0;
            const int r = vv & 127;
            const int dw = r & ((1 << p.bw_log2) - 1);
            const int dh = (r >> p.bw_log2) & ((1 << p.bh_log2) - 1);
            const int dt = (r >> (p.bw_log2 + p.bh_log2)) & ((1 << p.bt_log2) - 1);
            const int dn = r >> (p.bw_log2 + p.bh_log2 + p.bt_log2);
            const long long vox = (((long long)(tc.n0 + dn) * p.OT + tc.t0 + dt) * p.OH + tc.h0 + dh) * p.OW + tc.w0 + dw;
            const uint4 u = *reinterpret_cast<const uint4*>(buf + j * 256 + (lane & 15) * 16);
            *reinterpret_cast<uint4*>(outp + vox * p.ldo + n_tile * 128 + (lane & 15) * 8) = u;
          }
        }
        fl_s += st_s;
        fl_ss += st_ss;
      }
      for (int ms = 0; ms < (p.swap ? 0 : p.m_sub); ++ms) {
      const TileCoord tc = decode_m_tile(p, m_super * p.m_sub + ms);
      const int dw = row & ((1 << p.bw_log2) - 1);
      const int dh = (row >> p.bw_log2) & ((1 << p.bh_log2) - 1);
      const int dt = (row >> (p.bw_log2 + p.bh_log2)) & ((1 << p.bt_log2) - 1);
      const int dn = row >> (p.bw_log2 + p.bh_log2 + p.bt_log2);
      const int vn = tc.n0 + dn, vt = tc.t0 + dt, vh = tc.h0 + dh, vw = tc.w0 + dw;
      const bool row_ok = vn < p.N && vt < p.T && vh < p.H && vw < p.W;  // partial boxes: masked store
      // (generalised store position: a strided data gradient writes one residue class of the input grid per launch)
      const long long vox = (((long long)vn * p.OT + vt * p.om[0] + p.oo[0]) * p.OH + vh * p.om[1] + p.oo[1]) * p.OW +
                            vw * p.om[2] + p.oo[2];
      const int col0 = n_tile * p.block_n;

      const uint32_t t_addr = tmem_base + acc * kAccStride + ms * p.block_n + ((uint32_t)(q * 32) << 16);

      if (p.splits > 1) {
        // split-K: add this item's partial sums into the fp32 workspace (bias / cast happen in the finish pass)
        for (int c = 0; c < p.block_n; c += 32) {
          if (col0 + c >= p.n_out) break;
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c, v);
          tmem_ld_wait();
          if (row_ok && p.dbg != 1) {
            if (p.ws_slab) {
              // own slab: plain stores (each thread fills one 128-byte line of its row per chunk)
              float* dst = p.ws + (long long)(item - tile * p.splits) * p.ws_slab + vox * p.ldo + col0 + c;
              if (p.vec_ok && col0 + c + 32 <= p.n_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  __stcg(reinterpret_cast<float4*>(dst + j),
                         make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3])));
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + c + j < p.n_out) dst[j] = __uint_as_float(v[j]);
              }
            } else {
              float* dst = p.ws + vox * p.ldo + col0 + c;
              if (p.vec_ok && col0 + c + 32 <= p.n_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                               "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                               "f"(__uint_as_float(v[j + 3]))
                               : "memory");
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + c + j < p.n_out) atomicAdd(dst + j, __uint_as_float(v[j]));
              }
            }
          }
        }
        continue;
      }

      if (p.fast_store) {
        // bf16 output, whole 64-column chunks: registers -> (bias) -> bf16 -> swizzled smem staging -> the warp
        // writes 4 full 128-byte row segments per instruction (the row-per-thread TMEM layout would otherwise
        // scatter 16-byte pieces over 32 different lines per store).
        if (ms == 0) {
          asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's bias readers are done
          for (int j = threadIdx.x - 64; j < p.block_n; j += 128) {
            const int col = col0 + j;
            float b = 0.f;
            if (col < p.n_out) {
              if (p.bias0) b += __ldg(p.bias0 + col);
              if (p.bias1) b += __ldg(p.bias1 + col);
            }
            bias_s[j] = b;
            if (p.red_S) {
              coef_s[j] = col < p.n_out ? __ldg(p.red_A + (long long)tc.n0 * p.n_out + col) : 0.f;
              coef_s[256 + j] = col < p.n_out ? __ldg(p.red_B + (long long)tc.n0 * p.n_out + col) : 0.f;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        float st_s = 0.f, st_ss = 0.f;
        uint8_t* my_stage = stage_s + (warp - 2) * 4096;
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
        for (int c = 0; c < p.block_n; c += 64) {
          if (col0 + c >= p.n_out) break;  // partial last N tile (n_out % 64 == 0, so chunks are all-or-nothing)
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(t_addr + c, v0);
          tmem_ld_32x32(t_addr + c + 32, v1);
          tmem_ld_wait();
          if (p.residual && row_ok) {  // fp32 add before the single bf16 rounding
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + vox * p.ldo + col0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 u = __ldg(rp + j);
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hh[e]);
                uint32_t* dst = (j < 4) ? &v0[8 * j + 2 * e] : &v1[8 * (j - 4) + 2 * e];
                dst[0] = __float_as_uint(__uint_as_float(dst[0]) + f.x);
                dst[1] = __float_as_uint(__uint_as_float(dst[1]) + f.y);
              }
            }
          }
          // fold the bias in: from here on v0 / v1 are the final fp32 outputs of this row
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v0[j] = __float_as_uint(__uint_as_float(v0[j]) + bias_s[c + j]);
            v1[j] = __float_as_uint(__uint_as_float(v1[j]) + bias_s[c + 32 + j]);
          }
          if (p.gn_sums || p.red_S) {
            if (p.gn_sums && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float r0 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v0[j])));
                const float r1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v1[j])));
                st_s += r0 + r1;
                st_ss = fmaf(r0, r0, fmaf(r1, r1, st_ss));
              }
            }
            if (p.red_S) {
              float d1[64], d2[64];
              const uint4* xp = reinterpret_cast<const uint4*>(p.red_x + vox * p.ldo + col0 + c);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint4 u = make_uint4(0, 0, 0, 0);
                if (row_ok) u = __ldg(xp + j);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 xf = __bfloat1622float2(hh[e]);
#pragma unroll
                  for (int q2 = 0; q2 < 2; ++q2) {
                    const int idx = 8 * j + 2 * e + q2;
                    const float xv = q2 ? xf.y : xf.x;
                    const uint32_t raw = idx < 32 ? v0[idx] : v1[idx - 32];
                    float dv = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw)));  // what bwd_apply will read
                    if (p.red_act == 1) {
                      const float pre = fmaf(xv, coef_s[c + idx], coef_s[256 + c + idx]);
                      const float sg = 1.f / (1.f + __expf(-pre));
                      dv *= sg * (1.f + pre * (1.f - sg));
                    } else if (p.red_act >= 2) {  // LeakyReLU(0.01) / ReLU (act codes of norm_act.cu)
                      const float pre = fmaf(xv, coef_s[c + idx], coef_s[256 + c + idx]);
                      if (pre <= 0.f) dv *= (p.red_act == 2 ? 0.01f : 0.f);
                    }
                    if (!row_ok) dv = 0.f;
                    d1[idx] = dv;
                    d2[idx] = dv * xv;
                  }
                }
              }
              // warp transpose-reduce: afterwards lane l holds the 32-row sums of columns l and l+32
#pragma unroll
              for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                  float* arr = (w == 0 ? d1 : d2) + half * 32;
#pragma unroll
                  for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int j = 0; j < off; ++j) {
                      const float send = up ? arr[j] : arr[j + off];
                      const float keep = up ? arr[j + off] : arr[j];
                      arr[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                  }
                }
              }
              atomicAdd(&red_s[2 * (c + lane)], d1[0]);
              atomicAdd(&red_s[2 * (c + lane) + 1], d2[0]);
              atomicAdd(&red_s[2 * (c + 32 + lane)], d1[32]);
              atomicAdd(&red_s[2 * (c + 32 + lane) + 1], d2[32]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v0[8 * j + 0]), __uint_as_float(v0[8 * j + 1]));
            u.y = pack_bf16x2(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3]));
            u.z = pack_bf16x2(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5]));
            u.w = pack_bf16x2(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7]));
            *reinterpret_cast<uint4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = u;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v1[8 * j + 0]), __uint_as_float(v1[8 * j + 1]));
            u.y = pack_bf16x2(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3]));
            u.z = pack_bf16x2(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5]));
            u.w = pack_bf16x2(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7]));
            *reinterpret_cast<uint4*>(my_stage + lane * 128 + (((j + 4) ^ (lane & 7)) << 4)) = u;
          }
          __syncwarp();
          const int chunk = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const long long rvox = __shfl_sync(0xffffffffu, vox, r);
            const int rok = __shfl_sync(0xffffffffu, (int)row_ok, r);
            const uint4 u = *reinterpret_cast<const uint4*>(my_stage + r * 128 + ((chunk ^ (r & 7)) << 4));
            if (rok) *reinterpret_cast<uint4*>(outp + rvox * p.ldo + col0 + c + chunk * 8) = u;
          }
          __syncwarp();
        }
        fl_s += st_s;
        fl_ss += st_ss;
        continue;
      }

      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        if (p.block_n >= 32) {
          tmem_ld_32x32(t_addr + c, v);
        } else {
          uint32_t v16[16];
          tmem_ld_32x16(t_addr + c, v16);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0;
        }
        tmem_ld_wait();
        const int cbase = col0 + c;
        if (cbase >= p.n_out || !row_ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b = 0.f;
          const int col = cbase + j;
          if (col < p.n_out) {
            if (p.bias0) b += __ldg(p.bias0 + col);
            if (p.bias1) b += __ldg(p.bias1 + col);
          }
          f[j] = __uint_as_float(v[j]) + b;
          if (p.residual && col < p.n_out && row_ok) f[j] += __bfloat162float(p.residual[vox * p.ldo + col]);
        }
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + vox * p.ldo + cbase;
          if (p.vec_ok && cbase + 32 <= p.n_out) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (cbase + j < p.n_out) o[j] = f[j];
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * p.ldo + cbase;
          if (p.vec_ok && cbase + 32 <= p.n_out) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(f[j], f[j + 1]);
              u.y = pack_bf16x2(f[j + 2], f[j + 3]);
              u.z = pack_bf16x2(f[j + 4], f[j + 5]);
              u.w = pack_bf16x2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (cbase + j < p.n_out) o[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
      }  // m_sub
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (p.splits > 1 && p.tile_ctr) {
        // ---- fused split-K finish: the last of the `splits` items of this tile to arrive reduces nothing more — all
        // partial sums are already in `ws` (L2 reductions) — it reads the tile back, adds the bias, rounds, stores, emits
        // the GroupNorm sums and zeroes the tile for the next launch. Replaces a memset, a finish launch and a stats pass.
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et0 = threadIdx.x - 64;
        if (et0 == 0) {
          const unsigned int old = atomicAdd(p.tile_ctr + tile, 1u);
          const int last = old == (unsigned int)(p.splits - 1);
          if (last) p.tile_ctr[tile] = 0u;
          flag_s[0] = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (flag_s[0]) {
          __threadfence();
          float fs = 0.f, fss = 0.f;
          int n_of_tile = 0;
          // a warp walks its 32 rows one at a time, its lanes spread over the row's columns (8 per lane): every access is a
          // coalesced 1 KB row segment (the first version gave each thread a whole row: 32 scattered 32-byte pieces per
          // warp instruction made the finishing CTA 50 us slower than the separate finish launch it replaced)
          for (int ms = 0; ms < p.m_sub; ++ms) {
            const TileCoord tc = decode_m_tile(p, m_super * p.m_sub + ms);
            n_of_tile = tc.n0;
            const int col0 = n_tile * p.block_n;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              const int r = q * 32 + rr;
              const int dw = r & ((1 << p.bw_log2) - 1);
              const int dh = (r >> p.bw_log2) & ((1 << p.bh_log2) - 1);
              const int dt = (r >> (p.bw_log2 + p.bh_log2)) & ((1 << p.bt_log2) - 1);
              const int dn = r >> (p.bw_log2 + p.bh_log2 + p.bt_log2);
              const int vn = tc.n0 + dn, vt = tc.t0 + dt, vh = tc.h0 + dh, vw = tc.w0 + dw;
              if (!(vn < p.N && vt < p.T && vh < p.H && vw < p.W)) continue;
              const long long vox = (((long long)vn * p.OT + vt) * p.OH + vh) * p.OW + vw;
              for (int c = lane * 8; c < p.block_n; c += 256) {
                const int col = col0 + c;
                if (col >= p.n_out) break;
                float* wp = p.ws + vox * p.ldo + col;
                float f[8];
                if (p.vec_ok && col + 8 <= p.n_out) {
                  const float4 a = __ldcg(reinterpret_cast<const float4*>(wp)), b = __ldcg(reinterpret_cast<const float4*>(wp) + 1);
                  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                  __stcg(reinterpret_cast<float4*>(wp), make_float4(0.f, 0.f, 0.f, 0.f));
                  __stcg(reinterpret_cast<float4*>(wp) + 1, make_float4(0.f, 0.f, 0.f, 0.f));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    f[j] = (col + j < p.n_out) ? __ldcg(wp + j) : 0.f;
                    if (col + j < p.n_out) __stcg(wp + j, 0.f);
                  }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (col + j < p.n_out) {
                    if (p.bias0) f[j] += __ldg(p.bias0 + col + j);
                    if (p.bias1) f[j] += __ldg(p.bias1 + col + j);
                  }
                }
                if (p.out_f32) {
                  float* o = reinterpret_cast<float*>(p.out) + vox * p.ldo + col;
                  if (p.vec_ok && col + 8 <= p.n_out) {
                    *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
                  } else {
                    for (int j = 0; j < 8; ++j)
                      if (col + j < p.n_out) o[j] = f[j];
                  }
                } else {
                  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * p.ldo + col;
                  if (p.gn_sums) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (col + j < p.n_out) {
                        const float rv = __bfloat162float(__float2bfloat16_rn(f[j]));
                        fs += rv;
                        fss = fmaf(rv, rv, fss);
                      }
                  }
                  if (p.vec_ok && col + 8 <= p.n_out) {
                    uint4 u;
                    u.x = pack_bf16x2(f[0], f[1]);
                    u.y = pack_bf16x2(f[2], f[3]);
                    u.z = pack_bf16x2(f[4], f[5]);
                    u.w = pack_bf16x2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(o) = u;
                  } else {
                    for (int j = 0; j < 8; ++j)
                      if (col + j < p.n_out) o[j] = __float2bfloat16_rn(f[j]);
                  }
                }
              }
            }
          }
          if (p.gn_sums) {
            for (int o = 16; o > 0; o >>= 1) {
              fs += __shfl_xor_sync(0xffffffffu, fs, o);
              fss += __shfl_xor_sync(0xffffffffu, fss, o);
            }
            if (lane == 0) {
              atomicAdd(&p.gn_sums[(long long)n_of_tile * 2], (double)fs);
              atomicAdd(&p.gn_sums[(long long)n_of_tile * 2 + 1], (double)fss);
            }
          }
        }
      }
      if (p.fast_store && p.splits == 1 && (p.gn_sums || p.red_S)) {
        // flush this tile's fused reductions (all rows of a CTA tile belong to one sample: host-checked)
        const TileCoord tcf = decode_m_tile(p, m_super * p.m_sub);
        if (p.gn_sums) {
          for (int o = 16; o > 0; o >>= 1) {
            fl_s += __shfl_xor_sync(0xffffffffu, fl_s, o);
            fl_ss += __shfl_xor_sync(0xffffffffu, fl_ss, o);
          }
          if (lane == 0) {
            atomicAdd(&stat_s[0], (double)fl_s);
            atomicAdd(&stat_s[1], (double)fl_ss);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = threadIdx.x - 64;
        if (p.red_S) {
          const int col0f = n_tile * p.block_n;
          for (int j = et; j < 2 * p.block_n; j += 128) {
            const int col = col0f + (j >> 1);
            if (col < p.n_out) atomicAdd(&p.red_S[((long long)tcf.n0 * p.n_out + col) * 2 + (j & 1)], red_s[j]);
            red_s[j] = 0.f;
          }
        }
        if (p.gn_sums && et == 0) {
          atomicAdd(&p.gn_sums[(long long)tcf.n0 * 2], stat_s[0]);
          atomicAdd(&p.gn_sums[(long long)tcf.n0 * 2 + 1], stat_s[1]);
          stat_s[0] = 0.0;
          stat_s[1] = 0.0;
        }
        fl_s = fl_ss = 0.f;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// split-K finish: out = cast(ws + bias0 + bias1)
// Finish pass of a split-K launch: out = cast(sum of the `nslab` partial slabs + bias), and — optionally — the per-sample
// GroupNorm(1, C) sums of the bf16-rounded result (what og_gn_stats would compute in another pass over `out`).
// grid = (chunks, samples); `per_sample` = T*H*W*n_out elements.
template <int VEC>
__global__ void __launch_bounds__(256)
    og_splitk_finish_kernel(const float* __restrict__ ws, long long slab, int nslab, const float* __restrict__ bias0,
                            const float* __restrict__ bias1, void* __restrict__ out, int out_f32, int n_out,
                            long long per_sample, double* __restrict__ gn_sums) {
  const long long base = (long long)blockIdx.y * per_sample;
  float s = 0.f, ss = 0.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < per_sample;
       i += (long long)gridDim.x * blockDim.x * VEC) {
    float v[VEC];
    if (VEC == 4) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(ws + base + i));
      v[0] = a.x; v[1 % VEC] = a.y; v[2 % VEC] = a.z; v[3 % VEC] = a.w;
      for (int k = 1; k < nslab; ++k) {
        const float4 b = __ldcg(reinterpret_cast<const float4*>(ws + (long long)k * slab + base + i));
        v[0] += b.x; v[1 % VEC] += b.y; v[2 % VEC] += b.z; v[3 % VEC] += b.w;
      }
    } else {
      v[0] = __ldcg(ws + base + i);
      for (int k = 1; k < nslab; ++k) v[0] += __ldcg(ws + (long long)k * slab + base + i);
    }
    const int col = (int)(i % n_out);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (bias0) v[e] += __ldg(bias0 + col + e);
      if (bias1) v[e] += __ldg(bias1 + col + e);
    }
    if (out_f32) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) reinterpret_cast<float*>(out)[base + i + e] = v[e];
    } else {
      __nv_bfloat16 h[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        h[e] = __float2bfloat16_rn(v[e]);
        const float r = __bfloat162float(h[e]);
        s += r;
        ss = fmaf(r, r, ss);
      }
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + base + i;
      if (VEC == 4)
        *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(h);
      else
        o[0] = h[0];
    }
  }
  if (gn_sums) {
    __shared__ float red[2][8];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s;
      red[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        a += (double)red[0][w];
        b += (double)red[1][w];
      }
      atomicAdd(gn_sums + 2 * blockIdx.y, a);
      atomicAdd(gn_sums + 2 * blockIdx.y + 1, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int pick_block_n(int n_out, int num_m_tiles, bool mn_major) {
  // The widest N tile that covers n_out (<= 256). Measured on B200: fewer, fatter tiles beat more, thinner ones
  // even when they leave SMs idle (a 128x256 tile streams 96 B/MMA-clk, a 128x64 one 192 B/MMA-clk, and the
  // TMA-latency x smem-capacity product caps what one SM can stream); low tile counts are handled by split-K.
  int bn = 16;
  while (bn < n_out && bn < 256) bn *= 2;
  if (mn_major && bn < 64) bn = 64;
  (void)num_m_tiles;
  return bn;
}

struct IgemmLaunch {
  const void* a0 = nullptr;   // segment-0 input tensor (bf16 NDHWC), c0 channels, spatial extents aT x aH x aW
  int c0 = 0;
  int aT = 0, aH = 0, aW = 0;
  int astride[3] = {1, 1, 1};  // element stride of the A boxes along (t, h, w): forward strided convolution
  const void* a1 = nullptr;    // optional segment-1 tensor (the fused 1x1x1 shortcut), same extents as the output grid
  int c1 = 0;
  IgemmSeg segs[2];
  int nseg = 1;
  const void* w = nullptr;
  int ldw = 0, k_off = 0, b_mn_major = 0, b_rows = 0, b_ntaps = 0;
  const float* bias0 = nullptr;
  const float* bias1 = nullptr;
  const void* residual = nullptr;
  void* out = nullptr;
  int out_f32 = 0;
  int N = 0, T = 0, H = 0, W = 0;  // the M grid (voxels this launch computes)
  int OT = 0, OH = 0, OW = 0;      // output tensor extents (0 = same as the M grid)
  int om[3] = {1, 1, 1}, oo[3] = {0, 0, 0};
  int n_out = 0;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  double* gn_sums = nullptr;
  const void* red_x = nullptr;
  const float* red_A = nullptr;
  const float* red_B = nullptr;
  int red_act = 0;
  float* red_S = nullptr;
};

// workspaces prepared by og_workspace_init (all zero, magic in the tail): only those may use the in-kernel split-K finish
static std::mutex g_ws_mutex;
static std::vector<std::pair<const void*, size_t>> g_ws_prepared;
static bool ws_prepared(const void* p, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (const auto& e : g_ws_prepared)
    if (e.first == p && e.second == bytes) return true;
  return false;
}
static constexpr size_t kWsTail = 4096;   // [0,256): scheduler state; [256,4096): per-tile arrival counters

static IgemmSeg make_seg(int cin_blocks, int kt, int kh, int kw, int pt, int ph, int pw, int sgn) {
  IgemmSeg g;
  g.cin_blocks = cin_blocks;
  const int k[3] = {kt, kh, kw}, pd[3] = {pt, ph, pw};
  for (int d = 0; d < 3; ++d) {
    g.n[d] = k[d];
    g.tap0[d] = 0;
    g.tstep[d] = 1;
    g.sh0[d] = -sgn * pd[d];
    g.shstep[d] = sgn;
  }
  g.kh = kh;
  g.kw = kw;
  return g;
}

static int launch_igemm(const IgemmLaunch& L, cudaStream_t stream) {
  const int N = L.N, T = L.T, H = L.H, W = L.W, n_out = L.n_out;
  OG_REQUIRE(N > 0 && T > 0 && H > 0 && W > 0 && n_out > 0, "conv3d: empty problem");
  const bool plain = (L.OT == 0) && L.astride[0] == 1 && L.astride[1] == 1 && L.astride[2] == 1;
  int bw, bh, bt, bn;
  choose_voxel_box(kBlockM, N, T, H, W, &bw, &bh, &bt, &bn);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = L.nseg;
  p.num_kb = 0;
  for (int s = 0; s < L.nseg; ++s) {
    p.seg[s] = L.segs[s];
    p.num_kb += L.segs[s].cin_blocks * L.segs[s].n[0] * L.segs[s].n[1] * L.segs[s].n[2];
  }
  OG_REQUIRE(p.num_kb > 0, "conv3d: empty reduction");
  for (int d = 0; d < 3; ++d) {
    p.sx[d] = L.astride[d];
    p.om[d] = L.om[d];
    p.oo[d] = L.oo[d];
  }
  p.OT = L.OT ? L.OT : T;
  p.OH = L.OH ? L.OH : H;
  p.OW = L.OW ? L.OW : W;
  p.b_mn_major = L.b_mn_major;
  p.bw_log2 = ilog2(bw);
  p.bh_log2 = ilog2(bh);
  p.bt_log2 = ilog2(bt);
  p.bn_log2 = ilog2(bn);
  p.tiles_w = (W + bw - 1) / bw;
  p.tiles_h = (H + bh - 1) / bh;
  p.tiles_t = (T + bt - 1) / bt;
  p.N = N;
  p.T = T;
  p.H = H;
  p.W = W;
  p.num_m_tiles = ((N + bn - 1) / bn) * p.tiles_w * p.tiles_h * p.tiles_t;
  p.block_n = pick_block_n(n_out, p.num_m_tiles, L.b_mn_major != 0);
  if (const char* e = getenv("OG_IGEMM_BN")) {  // tuning experiments only
    const int v = atoi(e);
    if (v >= 16 && v <= 256 && (!L.b_mn_major || v >= 64)) p.block_n = v;
  }
  p.num_n_tiles = (n_out + p.block_n - 1) / p.block_n;
  p.n_out = n_out;
  p.ldo = n_out;
  p.out = L.out;
  p.out_f32 = L.out_f32;
  p.vec_ok = L.out_f32 ? (n_out % 4 == 0) : (n_out % 8 == 0);
  p.bias0 = L.bias0;
  p.bias1 = L.bias1;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(L.residual);
  // TMA latency x smem capacity bounds the bytes/clk one SM can stream; sharing each B stage between two
  // 128-row M sub-tiles keeps the demand at (32+16) KB per 512 MMA-clk (same as a 128x256 tile).
  p.m_sub = (p.block_n <= 128 && p.num_m_tiles >= 2 * num_sms()) ? 2 : 1;
  if (const char* e = getenv("OG_IGEMM_MSUB")) {
    const int v = atoi(e);
    if ((v == 1 || v == 2) && v * p.block_n <= 256) p.m_sub = v;
  }
  const int stage_bytes = p.m_sub * kABytes + p.block_n * kBlockK * 2;
  p.fast_store = (!L.out_f32 && n_out % 64 == 0 && p.block_n % 64 == 0) ? 1 : 0;
  // dynamic tile scheduling: state lives in the last 256 bytes of the caller's workspace (og_workspace_init)
  p.sched = nullptr;
  size_t ws_usable = L.workspace_bytes;
  unsigned int* tile_ctr = nullptr;
  if (L.workspace && L.workspace_bytes >= 2 * kWsTail) {
    const size_t off = (L.workspace_bytes - kWsTail) & ~(size_t)255;
    ws_usable = off;
    // in-kernel split-K finish: OFF by default. Measured on one B200 (profiles/r02p_*): it removes 98 launches per step
    // (memset + finish + statistics pass) yet the graphed step is 64.0-64.7 ms with it and 62.95 ms without — the last
    // arriver's fence + read-back sits on the tail of every split launch, while the three small launches it replaces cost
    // ~3 us each inside a graph. OG_SPLITK_FUSED=1 enables it.
    static const bool fused_finish_on = [] {
      const char* e = getenv("OG_SPLITK_FUSED");
      return e && atoi(e) == 1;
    }();
    if (fused_finish_on && ws_prepared(L.workspace, L.workspace_bytes))
      tile_ctr = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(L.workspace) + off + 256);
    // OFF by default: measured on 2 x B200 (profiles/r02k_*), the graphed data-parallel step takes 68.7 ms with either
    // assignment and the single-GPU step is within noise (64.9 dynamic vs 64.6 static) — the cost of overlapping the
    // all-reduce turned out to be power, not SM residency. OG_IGEMM_DYNAMIC=1 enables it.
    static const bool dyn_on = [] {
      const char* e = getenv("OG_IGEMM_DYNAMIC");
      return e && atoi(e) == 1;
    }();
    if (dyn_on) p.sched = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(L.workspace) + off);
  }
  // split-K when the tiles cannot fill the machine and each has a long K loop
  p.splits = 1;
  p.ws = nullptr;
  p.ws_slab = 0;
  p.dbg = 0;
  if (const char* e = getenv("OG_IGEMM_DBG")) p.dbg = atoi(e);
  if (plain) {
    const long long tiles = (long long)((p.num_m_tiles + p.m_sub - 1) / p.m_sub) * p.num_n_tiles;
    const size_t need = (size_t)N * T * H * W * n_out * sizeof(float);
    if (tiles * 2 <= num_sms() && p.num_kb >= 32 && L.workspace && ws_usable >= need && !L.residual) {
      int sp = (int)(num_sms() / tiles);
      if (sp > p.num_kb / 8) sp = p.num_kb / 8;
      if (sp > 16) sp = 16;
      if (sp >= 2) {
        p.splits = sp;
        p.ws = reinterpret_cast<float*>(L.workspace);
        // one slab per split when the workspace holds them (plain stores, no memset); OG_SPLITK_SLABS=0: one zeroed slab
        // that the items reduce into with red.global.add (the round-1 form, kept for small workspaces)
        static const bool slabs_on = [] {
          const char* e = getenv("OG_SPLITK_SLABS");
          return !(e && atoi(e) == 0);
        }();
        if (tile_ctr && tiles <= (long long)((kWsTail - 256) / sizeof(unsigned int)))
          p.tile_ctr = tile_ctr;     // prepared workspace: `ws` is zero on entry and is left zero (fused finish)
        else if (slabs_on && ws_usable >= need * (size_t)sp)
          p.ws_slab = (long long)(need / sizeof(float));
        else
          OG_CHECK_CUDA(cudaMemsetAsync(L.workspace, 0, need, stream));
      }
    }
  }
  const int tail_bytes = 256 /*barriers*/ + 1024 /*bias*/ + 4 * 4096 /*store staging*/ + 4352 /*fused reductions*/;
  int stages = (227 * 1024 - 1024 /*align slack*/ - tail_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.num_stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + tail_bytes;

  CUtensorMap mapA0, mapA1, mapB;
  {
    // forward strided convolution: the box spans bw*sw input positions and TMA traverses it with element stride sw,
    // landing exactly the bw voxels the tile needs (the hardware zero-fills the out-of-range ones: the padding)
    const int c0 = L.c0;
    const int st = L.astride[0], sh = L.astride[1], sw = L.astride[2];
    uint64_t dims[5] = {(uint64_t)c0, (uint64_t)L.aW, (uint64_t)L.aH, (uint64_t)L.aT, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)c0 * 2, (uint64_t)L.aW * c0 * 2, (uint64_t)L.aH * L.aW * c0 * 2,
                       (uint64_t)L.aT * L.aH * L.aW * c0 * 2};
    uint32_t box[5] = {kBlockK, (uint32_t)(bw * sw), (uint32_t)(bh * sh), (uint32_t)(bt * st), (uint32_t)bn};
    uint32_t es[5] = {1, (uint32_t)sw, (uint32_t)sh, (uint32_t)st, 1};
    OG_REQUIRE(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "conv3d: strided box exceeds the TMA limit");
    int r = make_tmap_bf16(&mapA0, L.a0, 5, dims, str, box, es);
    if (r != OG_OK) return r;
  }
  if (L.nseg > 1) {
    const int c1 = L.c1;
    uint64_t dims[5] = {(uint64_t)c1, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)c1 * 2, (uint64_t)W * c1 * 2, (uint64_t)H * W * c1 * 2, (uint64_t)T * H * W * c1 * 2};
    uint32_t box[5] = {kBlockK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bt, (uint32_t)bn};
    int r = make_tmap_bf16(&mapA1, L.a1, 5, dims, str, box);
    if (r != OG_OK) return r;
  } else {
    mapA1 = mapA0;
  }
  const __nv_bfloat16* wb = reinterpret_cast<const __nv_bfloat16*>(L.w) + L.k_off;
  if (!L.b_mn_major) {
    // rows = output channels, contiguous k
    uint64_t ktot = (uint64_t)p.num_kb * kBlockK;
    uint64_t dims[2] = {ktot, (uint64_t)n_out};
    uint64_t str[1] = {(uint64_t)L.ldw * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)p.block_n};
    int r = make_tmap_bf16(&mapB, wb, 2, dims, str, box);
    if (r != OG_OK) return r;
  } else {
    // w[co][tap][ci] viewed as (ci, tap, co): a (64,1,64) box lands as one [64 co rows][64 ci] panel
    uint64_t dims[3] = {(uint64_t)n_out, (uint64_t)L.b_ntaps, (uint64_t)L.b_rows};
    uint64_t str[2] = {(uint64_t)n_out * 2, (uint64_t)L.ldw * 2};
    uint32_t box[3] = {64, 1, 64};
    int r = make_tmap_bf16(&mapB, wb, 3, dims, str, box);
    if (r != OG_OK) return r;
  }

  static bool attr_set = false;
  if (!attr_set) {
    OG_CHECK_CUDA(cudaFuncSetAttribute(og_conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // operand swap for 128-wide Cout tiles (two voxel sub-tiles form the N = 256 operand): whole boxes only
  p.swap = (plain && p.block_n == 128 && p.m_sub == 2 && p.fast_store && p.splits == 1 && !L.residual && !L.red_S &&
            n_out % 128 == 0 && (p.num_m_tiles % 2 == 0) && W % bw == 0 && H % bh == 0 && T % bt == 0 && N % bn == 0)
               ? 1 : 0;
  if (const char* e = getenv("OG_IGEMM_SWAP")) {
    if (atoi(e) == 0) p.swap = 0;
  }
  // fused epilogue reductions need: staged bf16 stores, no split-K, every CTA tile inside one sample
  const int tiles_per_sample = p.tiles_w * p.tiles_h * p.tiles_t;
  const bool can_fuse = plain && p.fast_store && p.splits == 1 && bn == 1 && (tiles_per_sample % p.m_sub == 0) && n_out <= 65536;
  const bool can_fuse_splitk = plain && p.splits > 1 && p.tile_ctr && bn == 1 && !L.out_f32 && (tiles_per_sample % p.m_sub == 0);
  p.gn_sums = (L.gn_sums && (can_fuse || can_fuse_splitk)) ? L.gn_sums : nullptr;
  p.red_S = (L.red_S && can_fuse) ? L.red_S : nullptr;
  p.red_x = reinterpret_cast<const __nv_bfloat16*>(L.red_x);
  p.red_A = L.red_A;
  p.red_B = L.red_B;
  p.red_act = L.red_act;
  const int total_tiles = ((p.num_m_tiles + p.m_sub - 1) / p.m_sub) * p.num_n_tiles * p.splits;
  int grid = num_sms();
  if (grid > total_tiles) grid = total_tiles;
  og_conv_igemm_kernel<<<grid, kThreads, smem_bytes, stream>>>(mapA0, mapA1, mapB, p);
  OG_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  bool finish_did_stats = false;
  if (p.splits > 1 && !p.tile_ctr) {
    const long long per_sample = (long long)T * H * W * n_out;
    const int nslab = p.ws_slab ? p.splits : 1;
    finish_did_stats = L.gn_sums && !L.out_f32;
    double* gs = finish_did_stats ? L.gn_sums : nullptr;
    const int vec = (n_out % 4 == 0) ? 4 : 1;
    long long bx = (per_sample / vec + 255) / 256;
    const long long cap = ((long long)num_sms() * 8 + N - 1) / N;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)N);
    if (vec == 4)
      og_splitk_finish_kernel<4><<<grid, 256, 0, stream>>>(p.ws, p.ws_slab, nslab, L.bias0, L.bias1, L.out, L.out_f32,
                                                            n_out, per_sample, gs);
    else
      og_splitk_finish_kernel<1><<<grid, 256, 0, stream>>>(p.ws, p.ws_slab, nslab, L.bias0, L.bias1, L.out, L.out_f32,
                                                            n_out, per_sample, gs);
    OG_CHECK_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
  }
  // requested reductions that could not be fused: run the stand-alone passes on the stored output
  if (L.gn_sums && !p.gn_sums && !finish_did_stats) {
    OG_REQUIRE(!L.out_f32 && plain, "conv3d: GroupNorm statistics need a plain bf16 output");
    int r = og_gn_stats(L.out, N, (int64_t)T * H * W, n_out, 1, L.gn_sums, (og_stream_t)stream);
    if (r != OG_OK) return r;
  }
  if (L.red_S && !p.red_S) {
    OG_REQUIRE(!L.out_f32 && plain, "conv3d: fused backward reduction needs a plain bf16 output");
    int r = og_affine_act_bwd_reduce(L.out, L.red_x, L.red_A, L.red_B, L.red_act, L.red_S, N, (int64_t)T * H * W, n_out,
                                     (og_stream_t)stream);
    if (r != OG_OK) return r;
  }
  return OG_OK;
}

}  // namespace og

extern "C" int og_conv3d_fwd(const void* x0, int c0, int kt, int kh, int kw, int pt, int ph, int pw, const void* x1,
                             int c1, const void* w, int ldw, const float* bias0, const float* bias1,
                             const void* residual, void* out, int out_f32, int N, int T, int H, int W, int cout,
                             void* workspace, size_t workspace_bytes, double* gn_sums, og_stream_t stream) {
  using namespace og;
  OG_REQUIRE(x0 && w && out, "conv3d_fwd: null pointer");
  OG_REQUIRE(c0 > 0 && c0 % 64 == 0, "conv3d_fwd: c0=%d must be a positive multiple of 64 (use the im2col path)", c0);
  OG_REQUIRE(!x1 || (c1 > 0 && c1 % 64 == 0), "conv3d_fwd: c1=%d must be a multiple of 64", c1);
  OG_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && pt >= 0 && ph >= 0 && pw >= 0 && pt < kt && ph < kh && pw < kw,
             "conv3d_fwd: bad kernel/padding (%d,%d,%d)/(%d,%d,%d)", kt, kh, kw, pt, ph, pw);
  const int ktot = kt * kh * kw * c0 + (x1 ? c1 : 0);
  OG_REQUIRE(ldw >= ktot && ldw % 8 == 0, "conv3d_fwd: ldw=%d must be >= %d and a multiple of 8", ldw, ktot);
  IgemmLaunch L;
  L.a0 = x0; L.c0 = c0; L.aT = T; L.aH = H; L.aW = W;
  L.a1 = x1; L.c1 = c1;
  L.segs[0] = make_seg(c0 / 64, kt, kh, kw, pt, ph, pw, +1);
  L.segs[1] = make_seg(c1 / 64, 1, 1, 1, 0, 0, 0, +1);
  L.nseg = x1 ? 2 : 1;
  L.w = w; L.ldw = ldw;
  L.bias0 = bias0; L.bias1 = bias1; L.residual = residual;
  L.out = out; L.out_f32 = out_f32;
  L.N = N; L.T = T; L.H = H; L.W = W; L.n_out = cout;
  L.workspace = workspace; L.workspace_bytes = workspace_bytes;
  L.gn_sums = gn_sums;
  return launch_igemm(L, (cudaStream_t)stream);
}

extern "C" int og_conv3d_dgrad(const void* dy, int cout, int w_rows, const void* w, int ldw, int k_off, int kt, int kh,
                               int kw, int pt, int ph, int pw, void* dx, int dx_f32, int N, int T, int H, int W,
                               int cin, void* workspace, size_t workspace_bytes, const void* red_x,
                               const float* red_A, const float* red_B, int red_act, float* red_S,
                               og_stream_t stream) {
  using namespace og;
  OG_REQUIRE(dy && w && dx, "conv3d_dgrad: null pointer");
  OG_REQUIRE(cout > 0 && cout % 64 == 0, "conv3d_dgrad: cout=%d must be a multiple of 64", cout);
  OG_REQUIRE(cin > 0 && cin % 64 == 0, "conv3d_dgrad: cin=%d must be a multiple of 64", cin);
  OG_REQUIRE(k_off % 8 == 0 && ldw % 8 == 0, "conv3d_dgrad: k_off/ldw must be multiples of 8");
  OG_REQUIRE(w_rows > 0 && w_rows <= cout, "conv3d_dgrad: w_rows=%d must be in (0, cout]", w_rows);
  IgemmLaunch L;
  L.a0 = dy; L.c0 = cout; L.aT = T; L.aH = H; L.aW = W;
  L.segs[0] = make_seg(cout / 64, kt, kh, kw, pt, ph, pw, -1);
  L.nseg = 1;
  L.w = w; L.ldw = ldw; L.k_off = k_off; L.b_mn_major = 1; L.b_rows = w_rows; L.b_ntaps = kt * kh * kw;
  L.out = dx; L.out_f32 = dx_f32;
  L.N = N; L.T = T; L.H = H; L.W = W; L.n_out = cin;
  L.workspace = workspace; L.workspace_bytes = workspace_bytes;
  L.red_x = red_x; L.red_A = red_A; L.red_B = red_B; L.red_act = red_act; L.red_S = red_S;
  return launch_igemm(L, (cudaStream_t)stream);
}

// Prepares a workspace for og_conv3d_fwd / og_conv3d_dgrad: writes the dynamic tile scheduler's state {magic, 0, 0} into
// its last 256 bytes. Call once after allocating (or re-allocating) the buffer; the kernels reset the state themselves.
// A workspace that was not initialised (no magic) simply gets the static tile assignment.
extern "C" int og_workspace_init(void* workspace, size_t workspace_bytes, og_stream_t stream) {
  using namespace og;
  OG_REQUIRE(workspace && workspace_bytes >= 2 * kWsTail, "workspace_init: need a workspace of at least %zu bytes", 2 * kWsTail);
  const size_t off = (workspace_bytes - kWsTail) & ~(size_t)255;
  char* tail = reinterpret_cast<char*>(workspace) + off;
  OG_CHECK_CUDA(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));   // partial sums + counters: all zero
  static const unsigned int magic = kSchedMagic;
  OG_CHECK_CUDA(cudaMemcpyAsync(tail, &magic, sizeof(magic), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    bool found = false;
    for (auto& e : g_ws_prepared)
      if (e.first == workspace) {
        e.second = workspace_bytes;
        found = true;
      }
    if (!found) g_ws_prepared.emplace_back(workspace, workspace_bytes);
  }
  return OG_OK;
}

static int out_extent(int in, int pad_front, int pad_back, int k, int s) { return (in + pad_front + pad_back - k) / s + 1; }

// Strided CausalConv3d forward (SpaceTimeDownsample, genie/module/video.py:457-483; geometry video.py:154-164: time is
// padded at the FRONT only by pt = kt-1 + (1-st), space symmetrically by (k-1)//2), as the SAME implicit GEMM: the A
// box of a tap is a strided TMA box (element strides = the convolution strides). No im2col buffer.
extern "C" int og_conv3d_strided_fwd(const void* x, int cin, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph,
                                     int pw, const void* w, int ldw, const float* bias, void* out, int out_f32, int N, int T,
                                     int H, int W, int cout, og_stream_t stream) {
  using namespace og;
  OG_REQUIRE(x && w && out, "conv3d_strided_fwd: null pointer");
  OG_REQUIRE(cin > 0 && cin % 64 == 0, "conv3d_strided_fwd: cin=%d must be a positive multiple of 64", cin);
  OG_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && st >= 1 && sh >= 1 && sw >= 1 && st <= 8 && sh <= 8 && sw <= 8 && pt >= 0 &&
                 ph >= 0 && pw >= 0, "conv3d_strided_fwd: bad kernel / stride / padding");
  OG_REQUIRE(ldw >= kt * kh * kw * cin && ldw % 8 == 0, "conv3d_strided_fwd: bad ldw=%d", ldw);
  const int To = out_extent(T, pt, 0, kt, st), Ho = out_extent(H, ph, ph, kh, sh), Wo = out_extent(W, pw, pw, kw, sw);
  OG_REQUIRE(To >= 1 && Ho >= 1 && Wo >= 1, "conv3d_strided_fwd: empty output");
  IgemmLaunch L;
  L.a0 = x; L.c0 = cin; L.aT = T; L.aH = H; L.aW = W;
  L.astride[0] = st; L.astride[1] = sh; L.astride[2] = sw;
  L.segs[0] = make_seg(cin / 64, kt, kh, kw, pt, ph, pw, +1);
  L.nseg = 1;
  L.w = w; L.ldw = ldw;
  L.bias0 = bias;
  L.out = out; L.out_f32 = out_f32;
  L.N = N; L.T = To; L.H = Ho; L.W = Wo; L.n_out = cout;
  return launch_igemm(L, (cudaStream_t)stream);
}

// Data gradient of the strided convolution, without col2im: input position i receives dy[(i + pad - tap) / s] only from
// the taps with tap == (i + pad) mod s, so the input grid splits into st*sh*sw residue classes, each of which is a
// stride-1 implicit GEMM over dy with its own tap subset whose output rows are stored at stride s into dx.
// dy: bf16 [N,To,Ho,Wo,cout] (cout % 64 == 0, zero padded by the caller; w_rows real rows); dx: bf16 [N,T,H,W,cin].
extern "C" int og_conv3d_strided_dgrad(const void* dy, int cout, int w_rows, const void* w, int ldw, int kt, int kh, int kw,
                                       int st, int sh, int sw, int pt, int ph, int pw, void* dx, int N, int T, int H, int W,
                                       int cin, og_stream_t stream) {
  using namespace og;
  OG_REQUIRE(dy && w && dx, "conv3d_strided_dgrad: null pointer");
  OG_REQUIRE(cout > 0 && cout % 64 == 0 && cin > 0 && cin % 64 == 0, "conv3d_strided_dgrad: channels must be multiples of 64");
  OG_REQUIRE(w_rows > 0 && w_rows <= cout && ldw % 8 == 0, "conv3d_strided_dgrad: bad w_rows / ldw");
  const int To = out_extent(T, pt, 0, kt, st), Ho = out_extent(H, ph, ph, kh, sh), Wo = out_extent(W, pw, pw, kw, sw);
  OG_REQUIRE(To >= 1 && Ho >= 1 && Wo >= 1, "conv3d_strided_dgrad: empty output");
  const int k[3] = {kt, kh, kw}, s[3] = {st, sh, sw}, pd[3] = {pt, ph, pw}, in[3] = {T, H, W};
  bool any_empty = false;
  for (int d = 0; d < 3; ++d)
    if (k[d] < s[d]) any_empty = true;   // some residue classes have no tap: their rows of dx are zero
  if (any_empty) OG_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * T * H * W * cin * 2, (cudaStream_t)stream));
  for (int ct = 0; ct < st; ++ct)
    for (int ch = 0; ch < sh; ++ch)
      for (int cw = 0; cw < sw; ++cw) {
        const int cls[3] = {ct, ch, cw};
        IgemmLaunch L;
        L.segs[0].cin_blocks = cout / 64;
        L.segs[0].kh = kh;
        L.segs[0].kw = kw;
        int grid[3];
        bool empty = false;
        for (int d = 0; d < 3; ++d) {
          // class c: input positions i with (i + pad) mod s == c, i.e. i = s*a + r, r = (c - pad) mod s (>= 0)
          const int r = ((cls[d] - pd[d]) % s[d] + s[d]) % s[d];
          const int ntap = cls[d] < k[d] ? (k[d] - 1 - cls[d]) / s[d] + 1 : 0;
          grid[d] = in[d] > r ? (in[d] - r + s[d] - 1) / s[d] : 0;
          if (ntap == 0 || grid[d] == 0) empty = true;
          L.segs[0].n[d] = ntap;
          L.segs[0].tap0[d] = cls[d];
          L.segs[0].tstep[d] = s[d];
          L.segs[0].sh0[d] = (r + pd[d] - cls[d]) / s[d];   // exact: r + pad - c is a multiple of s
          L.segs[0].shstep[d] = -1;
          L.om[d] = s[d];
          L.oo[d] = r;
        }
        if (empty) continue;
        L.a0 = dy; L.c0 = cout; L.aT = To; L.aH = Ho; L.aW = Wo;
        L.nseg = 1;
        L.w = w; L.ldw = ldw; L.b_mn_major = 1; L.b_rows = w_rows; L.b_ntaps = kt * kh * kw;
        L.out = dx; L.out_f32 = 0;
        L.N = N; L.T = grid[0]; L.H = grid[1]; L.W = grid[2]; L.n_out = cin;
        L.OT = T; L.OH = H; L.OW = W;
        int rc = launch_igemm(L, (cudaStream_t)stream);
        if (rc != OG_OK) return rc;
      }
  return OG_OK;
}
