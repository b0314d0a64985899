/* opengenie_b200.h — C ABI of libopengenie_b200.so (sm_100a only).
 *
 * The drop-in boundary for open-genie's data-parallel hot path. The reference has no FFI of its own:
 * every entry point below replaces one ATen call site (or a fused run of them) in the reference's
 * Python modules; the file:line each one replaces is cited next to its declaration
 * (paths relative to the reference repo, myscience/open-genie @ 732b9f9).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer (device memory) and the stream.
 *   - activations are NDHWC ("channels-last-3d") bf16 unless stated otherwise: x[n][t][h][w][c].
 *   - convolution weights live in "packed" order w[cout][tap][cin] (tap = (it*kh + ih)*kw + iw), which
 *     is exactly the memory order of a torch (Cout,Cin,kt,kh,kw) tensor in channels_last_3d format.
 *   - every function returns OG_OK (0) or a negative og_status; og_last_error() gives the message.
 *     Nothing throws, nothing allocates device memory, nothing synchronises the stream.
 *   - there is NO CPU or library fallback: on a non-sm_100 device the launch fails with OG_ERR_CUDA.
 */
#ifndef OPENGENIE_B200_H_
#define OPENGENIE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum og_status {
  OG_OK = 0,
  OG_ERR_INVALID_ARGUMENT = -1,
  OG_ERR_UNSUPPORTED_SHAPE = -2,
  OG_ERR_CUDA = -3,
} og_status;

typedef void* og_stream_t; /* cudaStream_t */

/* Last error message of the calling thread ("" if none). */
const char* og_last_error(void);
/* Library/ABI version and the SM architecture it was compiled for (100). */
int og_abi_version(void);
int og_compiled_sm(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t og_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * conv3d — implicit GEMM on tcgen05 tensor cores (TMA-staged NDHWC tiles, fp32 accumulate in TMEM)
 * ---------------------------------------------------------------------------------------------- */

/* Forward 3-D convolution, stride 1, "same" output size, zero padding.
 * Replaces: F.pad + nn.Conv3d in CausalConv3d.forward (genie/module/video.py:178-192) with
 * pad_t_front = kt-1; nn.Conv3d(k=3,p=1) in VideoResidualBlock.main (video.py:609-629) with
 * pad_t_front = 1; the ST-block FFN conv (genie/module/attention.py:429-438, misc.py:94-97);
 * and, through the optional second segment, the always-present 1x1x1 shortcut conv plus the residual
 * add `main(x) + res(x)` (video.py:599-603, 648).
 *
 *   out[n,t,h,w,co] = bias0[co] + bias1[co]
 *                   + sum_{it,ih,iw,ci} x0[n, t+it-pt, h+ih-ph, w+iw-pw, ci] * w[co][(it,ih,iw)][ci]
 *                   + sum_{ci}          x1[n, t, h, w, ci]                  * w[co][kt*kh*kw*c0 + ci]
 *
 * x0: bf16 [N,T,H,W,c0], c0 % 64 == 0.  x1: bf16 [N,T,H,W,c1] or NULL (c1 % 64 == 0).
 * w : bf16 [cout][ldw], ldw >= kt*kh*kw*c0 + c1, ldw % 8 == 0.
 * out: [N,T,H,W,cout], bf16 or fp32 (out_f32 != 0). bias0/bias1: fp32 [cout] or NULL.
 * Reads outside [0,T)x[0,H)x[0,W) are zero (pad_mode='constant'). */
int og_conv3d_fwd(const void* x0, int c0, int kt, int kh, int kw, int pt, int ph, int pw, const void* x1, int c1,
                  const void* w, int ldw, const float* bias0, const float* bias1, void* out, int out_f32, int N,
                  int T, int H, int W, int cout, og_stream_t stream);

/* Data gradient of the same convolution (autograd's conv3d backward-input, reached from
 * video.py:192 / 609-629 / 599-603 during loss.backward()).
 *   dx[n,t,h,w,ci] = sum_{it,ih,iw,co} dy[n, t-(it-pt), h-(ih-ph), w-(iw-pw), co] * w[co][k_off + tap*cin + ci]
 * dy: bf16 [N,T,H,W,cout] (cout % 64 == 0; a narrower gradient is zero-padded by the caller and
 * w_rows <= cout gives the number of real weight rows); w as above (k_off selects the segment inside
 * a packed row, k_off % 8 == 0); dx: [N,T,H,W,cin] (cin % 64 == 0), bf16 or fp32. */
int og_conv3d_dgrad(const void* dy, int cout, int w_rows, const void* w, int ldw, int k_off, int kt, int kh, int kw,
                    int pt, int ph, int pw, void* dx, int dx_f32, int N, int T, int H, int W, int cin,
                    og_stream_t stream);

/* Weight gradient (autograd's conv3d backward-weight). ACCUMULATES into dw (caller zeroes it):
 *   dw[co][tap][ci] += sum_{n,t,h,w} dy[n,t,h,w,co] * x[n, t+it-pt, h+ih-ph, w+iw-pw, ci]
 * dy: bf16 [N,T,H,W,cout]; x: bf16 [N,T,H,W,cin] (cin % 64 == 0); dw: fp32, row stride ld_dw elements,
 * tap-major / channel-minor inside a row (the channels_last_3d order of the torch weight). */
int og_conv3d_wgrad(const void* dy, int cout, const void* x, int cin, float* dw, int64_t ld_dw, int kt, int kh,
                    int kw, int pt, int ph, int pw, int N, int T, int H, int W, og_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENGENIE_B200_H_ */
