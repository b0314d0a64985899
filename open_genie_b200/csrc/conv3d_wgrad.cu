// conv3d_wgrad.cu — weight gradient of the stride-1 3-D convolution on tcgen05 tensor cores.
//
//   dW[co][tap][ci] += sum_v dY[v][co] * X[v + tap - pad][ci]          (v = voxel (n,t,h,w))
//
// As a GEMM the reduction dimension is the voxel index, and both operands are stored with the
// NON-reduced dimension contiguous (channels-last), i.e. both are "MN-major" UMMA operands:
//   A tile = dY box  : [co/64 panel][64 voxel rows][64 co]   (TMA 5-D box, 128-byte swizzle)
//   B tile = X  box  : [ci/64 panel][64 voxel rows][64 ci]   (same box shifted by the tap; OOB -> 0)
// so no transposed copy of activations or gradients is ever made. Each filter tap owns its own fp32
// accumulator block in TMEM (up to 4 taps x 128 columns = all 512 columns); the dY tile of a k-step is
// loaded once and reused by every tap of the group. The voxel range is split across CTAs (split-K) and
// partial sums are combined with fp32 red.global.add into dW, which the caller zeroes.
// Optional fused bias gradient (og_conv3d_wgrad_bias): db[co] = sum_v dY[v][co] is the same GEMM against a column of
// ones — the CTAs of the LAST tap group (which has accumulator columns to spare whenever ntaps % 4 != 0 or Cin tiles are
// 64 wide) issue one extra N = 16 MMA per k-step against a constant all-ones tile in shared memory (a tile of ones is
// invariant under the 128-byte swizzle, so any valid descriptor reads it correctly); 4 of ~100 MMAs, no extra pass over dY.
//
// Replaces autograd's conv3d weight-gradient reached from genie/module/video.py:192,609-629,599-603 and
// genie/module/attention.py:429-438 during loss.backward().
#include <stdlib.h>

#include "og_host.cuh"
#include "og_ptx.cuh"

namespace og {
extern std::atomic<uint64_t> g_launches;

struct WgradParams {
  int kt, kh, kw, pt, ph, pw;
  int ntaps, taps_per_group, num_groups, splitk;
  int block_n;  // ci tile (64 or 128)
  int bw_log2, bh_log2, bt_log2, bn_log2;
  int tiles_w, tiles_h, tiles_t;
  int num_ksteps;  // 64-voxel boxes in the whole tensor
  int cin, cout;
  float* dw;
  long long ld_dw;
  int a_stages, b_stages;
  int sx_t, sx_h, sx_w;  // X box start = dY box start * stride + tap offset (strided convolution; 1 otherwise)
  int vec_ok;  // dw rows 16-byte aligned: vector reductions
  float* dbias;     // optional: += column sums of dY for output channels < n_bias (the convolution's bias gradient)
  int n_bias;
  int bias_col;     // TMEM column of the 16-wide ones-product block (last tap group only)
  int plain_store;  // OG_WGRAD_PLAIN_STORE=1: the ABI says "accumulates", so overwriting is opt-in (the Python side zeroes dw anyway)
  int dbg;     // timing experiments only (OG_WGRAD_DBG): 1 = pretend A is K-major, 2 = pretend B is K-major, 4 = skip epilogue
};

static constexpr int kWThreads = 192;
static constexpr int kVox = 64;                 // voxels (K rows) per k-step
static constexpr int kPanelBytes = kVox * 128;  // one 64-channel panel of a box: 8 KiB
static constexpr int kWABytes = 2 * kPanelBytes;  // 128 output channels
static constexpr int kWMaxStages = 12;

__global__ void __launch_bounds__(kWThreads, 1)
    og_conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX,
                         const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // One B stage holds a PAIR of taps back to back ([tap j panels][tap j+1 panels]); issued as a single
  // UMMA of N = 2*block_n whose accumulator columns are the two taps' blocks. That halves the smem bytes
  // the tensor core reads per MMA-clock for the A (dY) operand: 128x128 MMAs are smem-bandwidth bound.
  const int tap_bytes = (p.block_n / 64) * kPanelBytes;
  const int b_bytes = 2 * tap_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * kWABytes;
  uint8_t* smem_ones = smem_b + p.b_stages * b_bytes;   // 16 k-rows x 128 B of bf16 1.0 (bias-gradient operand), 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + 2048);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + kWMaxStages;
  uint64_t* full_b = bars + 2 * kWMaxStages;
  uint64_t* empty_b = bars + 3 * kWMaxStages;
  uint64_t* tmem_full = bars + 4 * kWMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kWMaxStages + 1);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128;
  const int ci0 = blockIdx.y * p.block_n;
  // tap group is the FAST index: the CTAs that stream the same voxel range (same split, different taps)
  // are co-scheduled, so dY / X tiles are fetched from DRAM once and hit in L2 for the other groups
  const int split = blockIdx.z / p.num_groups;
  const int group = blockIdx.z - split * p.num_groups;
  const int tap0 = group * p.taps_per_group;
  const int ntap = min(p.taps_per_group, p.ntaps - tap0);
  const bool do_bias = p.dbias != nullptr && blockIdx.y == 0 && group == p.num_groups - 1;
  // contiguous k-step range of this split
  const int ks_begin = (int)(((long long)p.num_ksteps * split) / p.splitk);
  const int ks_end = (int)(((long long)p.num_ksteps * (split + 1)) / p.splitk);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapDY);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (do_bias) {
    for (int i = threadIdx.x; i < 2048 / 4; i += kWThreads) reinterpret_cast<uint32_t*>(smem_ones)[i] = 0x3F803F80u;
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      // whole warp runs the loop converged; one elected lane issues (see elect_one() in og_ptx.cuh)
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      // The producer is ONE thread: keep its per-k-step instruction count tiny (no divisions in the loop).
      // tap offsets of this group, decoded once
      int off_w[4], off_h[4], off_t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tap = tap0 + (j < ntap ? j : 0);
        off_t[j] = tap / (p.kh * p.kw) - p.pt;
        off_h[j] = (tap / p.kw) % p.kh - p.ph;
        off_w[j] = tap % p.kw - p.pw;
      }
      // box coordinates of the first k-step, then advanced incrementally with carries
      const int per_sample = p.tiles_w * p.tiles_h * p.tiles_t;
      int tn = ks_begin / per_sample;
      int r = ks_begin - tn * per_sample;
      int tw = r % p.tiles_w;
      r /= p.tiles_w;
      int th = r % p.tiles_h;
      int tt = r / p.tiles_h;
      const int panels = p.block_n / 64;
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        const int n = tn << p.bn_log2, w0 = tw << p.bw_log2, h0 = th << p.bh_log2, t0 = tt << p.bt_log2;
        mbar_wait(&empty_a[sa], pha ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_a[sa], kWABytes);
          tma_load_5d(smem_a + sa * kWABytes, &mapDY, &full_a[sa], co0, w0, h0, t0, n);
          tma_load_5d(smem_a + sa * kWABytes + kPanelBytes, &mapDY, &full_a[sa], co0 + 64, w0, h0, t0, n);
        }
        __syncwarp();
        if (++sa == p.a_stages) {
          sa = 0;
          pha ^= 1;
        }
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (j < ntap) {
            const int nt = (ntap - j) < 2 ? 1 : 2;
            mbar_wait(&empty_b[sb], phb ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full_b[sb], (uint32_t)(nt * tap_bytes));
              uint8_t* dst = smem_b + sb * b_bytes;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (u < nt) {
                  for (int pp = 0; pp < panels; ++pp)
                    tma_load_5d(dst + u * tap_bytes + pp * kPanelBytes, &mapX, &full_b[sb], ci0 + pp * 64,
                                w0 * p.sx_w + off_w[j + u], h0 * p.sx_h + off_h[j + u], t0 * p.sx_t + off_t[j + u], n);
                }
              }
            }
            __syncwarp();
            if (++sb == p.b_stages) {
              sb = 0;
              phb ^= 1;
            }
          }
        }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            if (++tt == p.tiles_t) {
              tt = 0;
              ++tn;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      const uint32_t amn = (p.dbg & 1) ? 0u : 1u, bmn = (p.dbg & 2) ? 0u : 1u;
      const uint32_t idesc1 = umma_idesc_bf16(128, (uint32_t)p.block_n, amn, bmn);
      const uint32_t idesc2 = umma_idesc_bf16(128, (uint32_t)(2 * p.block_n), amn, bmn);
      const uint32_t idesc_ones = umma_idesc_bf16(128, 16u, amn, 1u);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        mbar_wait(&full_a[sa], pha);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + sa * kWABytes);
        for (int j = 0; j < ntap; j += 2) {
          mbar_wait(&full_b[sb], phb);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smem_b + sb * b_bytes);
          const uint32_t d_tmem = tmem_base + j * p.block_n;
          const uint32_t idesc = (ntap - j >= 2) ? idesc2 : idesc1;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kVox / 16; ++k) {
              // MN-major panels [panel][64 k-rows][128 B]: 16 k-rows = 2048 B, panel stride = 8192 B
              const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 2048, kPanelBytes, 1024);
              const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 2048, kPanelBytes, 1024);
              umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (ks > ks_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_b[sb]);
          }
          __syncwarp();
          if (++sb == p.b_stages) {
            sb = 0;
            phb ^= 1;
          }
        }
        if (do_bias && elect_one()) {   // db += dY^T . 1  (N = 16 columns of ones; every column holds the same sum)
          const uint32_t ones_addr = smem_u32(smem_ones);
#pragma unroll
          for (int k = 0; k < kVox / 16; ++k)
            umma_bf16_ss(tmem_base + p.bias_col, umma_smem_desc_sw128(a_addr + k * 2048, kPanelBytes, 1024),
                         umma_smem_desc_sw128(ones_addr, kPanelBytes, 1024), idesc_ones, (ks > ks_begin || k > 0) ? 1u : 0u);
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty_a[sa]);
        __syncwarp();
        if (++sa == p.a_stages) {
          sa = 0;
          pha ^= 1;
        }
      }
      if (elect_one()) umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    if (p.dbg & 8)
      mbar_wait(tmem_full, 0);
    else
      mbar_wait_relaxed(tmem_full, 0);
    tc_fence_after();
    for (int j = 0; j < ntap && !(p.dbg & 4); ++j) {
      const int tap = tap0 + j;
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + j * p.block_n + c + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.cout) {
          float* dst = p.dw + (long long)co * p.ld_dw + (long long)tap * p.cin + ci0 + c;
          if (p.vec_ok && ci0 + c + 32 <= p.cin && p.splitk == 1 && p.plain_store) {
            // one CTA owns this (co, tap, ci) block and the caller's buffer is freshly zeroed: plain vector stores
            // instead of 32 L2 reductions per thread and chunk (the low-resolution layers run with split-K = 1)
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              *reinterpret_cast<float4*>(dst + jj) = make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                                                                 __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
          } else if (p.vec_ok && ci0 + c + 32 <= p.cin) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + jj), "f"(__uint_as_float(v[jj])),
                           "f"(__uint_as_float(v[jj + 1])), "f"(__uint_as_float(v[jj + 2])),
                           "f"(__uint_as_float(v[jj + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (ci0 + c + jj < p.cin) atomicAdd(dst + jj, __uint_as_float(v[jj]));
          }
        }
      }
    }
    if (do_bias) {
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + p.bias_col + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      if (co < p.n_bias) atomicAdd(p.dbias + co, __uint_as_float(v[0]));
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace og

// N,T,H,W: the dY grid (= output voxels). Ti,Hi,Wi / st,sh,sw: extents of x and the convolution strides.
static int launch_wgrad(const void* dy, int cout, const void* x, int cin, float* dw, int64_t ld_dw, int kt, int kh, int kw,
                        int pt, int ph, int pw, int N, int T, int H, int W, int Ti, int Hi, int Wi, int st, int sh, int sw,
                        og_stream_t stream, float* dbias = nullptr, int n_bias = 0) {
  using namespace og;
  OG_REQUIRE(dy && x && dw, "conv3d_wgrad: null pointer");
  OG_REQUIRE(!dbias || (n_bias > 0 && n_bias <= cout), "conv3d_wgrad: n_bias=%d must be in 1..cout", n_bias);
  OG_REQUIRE(cin > 0 && cin % 64 == 0, "conv3d_wgrad: cin=%d must be a multiple of 64", cin);
  OG_REQUIRE(cout > 0 && cout % 8 == 0, "conv3d_wgrad: cout=%d must be a multiple of 8 (TMA row stride)", cout);
  OG_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && pt >= 0 && ph >= 0 && pw >= 0 && pt < kt && ph < kh && pw < kw,
             "conv3d_wgrad: bad kernel/padding");
  OG_REQUIRE(st >= 1 && sh >= 1 && sw >= 1 && st <= 8 && sh <= 8 && sw <= 8, "conv3d_wgrad: bad stride");
  int bw, bh, bt, bn;
  choose_voxel_box(kVox, N, T, H, W, &bw, &bh, &bt, &bn);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.kt = kt; p.kh = kh; p.kw = kw; p.pt = pt; p.ph = ph; p.pw = pw;
  p.sx_t = st; p.sx_h = sh; p.sx_w = sw;
  p.ntaps = kt * kh * kw;
  p.block_n = (cin % 128 == 0) ? 128 : 64;
  p.taps_per_group = 512 / p.block_n;
  if (p.taps_per_group > 4) p.taps_per_group = 4;  // smem budget of the B ring
  if (p.taps_per_group > p.ntaps) p.taps_per_group = p.ntaps;
  p.num_groups = (p.ntaps + p.taps_per_group - 1) / p.taps_per_group;
  p.bw_log2 = ilog2(bw); p.bh_log2 = ilog2(bh); p.bt_log2 = ilog2(bt); p.bn_log2 = ilog2(bn);
  p.tiles_w = (W + bw - 1) / bw; p.tiles_h = (H + bh - 1) / bh; p.tiles_t = (T + bt - 1) / bt;
  p.num_ksteps = ((N + bn - 1) / bn) * p.tiles_w * p.tiles_h * p.tiles_t;
  p.cin = cin; p.cout = cout; p.dw = dw; p.ld_dw = ld_dw;
  const int co_tiles = (cout + 127) / 128;
  const int ci_tiles = cin / p.block_n;
  const int base_ctas = co_tiles * ci_tiles * p.num_groups;
  // split-K factor: fill whole waves (a partial last wave costs a full CTA duration), keep >= 4 k-steps per
  // CTA so pipeline fill and the epilogue reductions stay amortised, at most ~3 waves.
  const int sms = num_sms();
  int max_split = p.num_ksteps / 4;
  if (max_split < 1) max_split = 1;
  if (max_split > 128) max_split = 128;
  int splitk = 1;
  double best = -1.0;
  for (int s = 1; s <= max_split; ++s) {
    const long long total = (long long)base_ctas * s;
    const long long waves = (total + sms - 1) / sms;
    if (waves > 3 && s > 1) break;
    const double eff = (double)total / (double)(waves * sms);
    // prefer higher wave efficiency; among equals prefer fewer splits (fewer reductions)
    if (eff > best + 1e-9) {
      best = eff;
      splitk = s;
    }
  }
  p.splitk = splitk;
  {
    const char* e = getenv("OG_WGRAD_DBG");
    p.dbg = e ? atoi(e) : 0;
    const char* ps = getenv("OG_WGRAD_PLAIN_STORE");
    p.plain_store = ps ? atoi(ps) : 0;
  }
  p.vec_ok = (ld_dw % 4 == 0) && (cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(dw) & 15) == 0);
  // fused bias gradient: the last tap group needs 16 spare accumulator columns
  bool colsum_after = false;
  if (dbias) {
    const int ntap_last = p.ntaps - (p.num_groups - 1) * p.taps_per_group;
    if (ntap_last * p.block_n + 16 <= 512 && !(p.dbg & 3)) {
      p.dbias = dbias;
      p.n_bias = n_bias;
      p.bias_col = ntap_last * p.block_n;
    } else {
      colsum_after = true;   // 4 full 128-wide tap blocks: no room — separate pass over dY
    }
  }
  const int b_bytes = 2 * (p.block_n / 64) * kPanelBytes;  // a pair of taps per stage
  p.a_stages = 3;
  p.b_stages = (int)((216 * 1024 - p.a_stages * kWABytes) / b_bytes);
  if (p.b_stages > kWMaxStages) p.b_stages = kWMaxStages;
  const size_t smem_bytes = (size_t)p.a_stages * kWABytes + (size_t)p.b_stages * b_bytes + 2048 /*ones*/ + 1024 + 512;

  CUtensorMap mapDY, mapX;
  {
    uint64_t dims[5] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)cout * 2, (uint64_t)W * cout * 2, (uint64_t)H * W * cout * 2,
                       (uint64_t)T * H * W * cout * 2};
    uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bt, (uint32_t)bn};
    int r = make_tmap_bf16(&mapDY, dy, 5, dims, str, box);
    if (r != OG_OK) return r;
  }
  {
    // strided convolution: the x box spans bw*sw positions traversed with element stride sw (see conv3d_igemm.cu)
    uint64_t dims[5] = {(uint64_t)cin, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)Ti, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)cin * 2, (uint64_t)Wi * cin * 2, (uint64_t)Hi * Wi * cin * 2,
                       (uint64_t)Ti * Hi * Wi * cin * 2};
    uint32_t box[5] = {64, (uint32_t)(bw * sw), (uint32_t)(bh * sh), (uint32_t)(bt * st), (uint32_t)bn};
    uint32_t es[5] = {1, (uint32_t)sw, (uint32_t)sh, (uint32_t)st, 1};
    OG_REQUIRE(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "conv3d_wgrad: strided box exceeds the TMA limit");
    int r = make_tmap_bf16(&mapX, x, 5, dims, str, box, es);
    if (r != OG_OK) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    OG_CHECK_CUDA(cudaFuncSetAttribute(og_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(co_tiles, ci_tiles, p.num_groups * p.splitk);  // z = split * num_groups + group
  og_conv_wgrad_kernel<<<grid, kWThreads, smem_bytes, (cudaStream_t)stream>>>(mapDY, mapX, p);
  OG_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  if (colsum_after) return og_colsum(dy, (int64_t)N * T * H * W, n_bias, cout, dbias, stream);
  return OG_OK;
}

extern "C" int og_conv3d_wgrad(const void* dy, int cout, const void* x, int cin, float* dw, int64_t ld_dw, int kt,
                               int kh, int kw, int pt, int ph, int pw, int N, int T, int H, int W,
                               og_stream_t stream) {
  return launch_wgrad(dy, cout, x, cin, dw, ld_dw, kt, kh, kw, pt, ph, pw, N, T, H, W, T, H, W, 1, 1, 1, stream);
}

// Weight gradient + bias gradient in one launch: dbias[c] += sum over voxels of dy[v][c] for c < n_bias (the bias
// gradient of the same nn.Conv3d); see the file header. Falls back to an og_colsum pass when the tiling has no spare
// accumulator columns.
extern "C" int og_conv3d_wgrad_bias(const void* dy, int cout, const void* x, int cin, float* dw, int64_t ld_dw, int kt,
                                    int kh, int kw, int pt, int ph, int pw, int N, int T, int H, int W, float* dbias,
                                    int n_bias, og_stream_t stream) {
  if (!dbias) {
    og::set_error("conv3d_wgrad_bias: dbias is NULL");
    return OG_ERR_INVALID_ARGUMENT;
  }
  return launch_wgrad(dy, cout, x, cin, dw, ld_dw, kt, kh, kw, pt, ph, pw, N, T, H, W, T, H, W, 1, 1, 1, stream, dbias, n_bias);
}

// Weight gradient of the strided CausalConv3d (SpaceTimeDownsample): dy on the OUTPUT grid, x on the input grid
// [N,T,H,W,cin]; x boxes are strided TMA boxes. Geometry as og_conv3d_strided_fwd.
extern "C" int og_conv3d_strided_wgrad(const void* dy, int cout, const void* x, int cin, float* dw, int64_t ld_dw, int kt,
                                       int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int N, int T, int H,
                                       int W, og_stream_t stream) {
  const int To = (T + pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  if (To < 1 || Ho < 1 || Wo < 1) {
    og::set_error("conv3d_strided_wgrad: empty output");
    return OG_ERR_INVALID_ARGUMENT;
  }
  return launch_wgrad(dy, cout, x, cin, dw, ld_dw, kt, kh, kw, pt, ph, pw, N, To, Ho, Wo, T, H, W, st, sh, sw, stream);
}
