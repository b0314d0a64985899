#!/usr/bin/env python
"""bench.py — VideoTokenizer train-step frames/sec @ 16x64x64 (BASELINE.json's metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU via torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port), rank 0 only

A "step" is one full training step of the MAGVIT2 VideoTokenizer (GAN / perceptual terms disabled — the
only configuration in which the reference runs offline) on one synthetic batch of
B x 3 x 16 x 64 x 64 video per GPU: forward, backward, (gradient all-reduce for N>1), fused AdamW.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, RES = 16, 64
# algorithmic conv FLOPs of one MAGVIT2 training step per clip (2506.1 GF forward, x3 for train)
CONV_GFLOP_FWD_PER_CLIP = 2506.1


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm (oracle port) on the host cores.  Test infrastructure used as a
# measured baseline only — never on the product path.
# --------------------------------------------------------------------------------------------------
def pick_cpu_threads():
    """MKLDNN conv3d on a big box is often FASTER with fewer threads than os.cpu_count() (cgroup quotas,
    oversubscription): time one representative conv at a few thread counts (~1 s each) and keep the best."""
    import torch.nn.functional as F
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, 128, avail) if c <= avail})
    x = torch.randn(1, 128, 8, 64, 64)
    w = torch.randn(128, 128, 3, 3, 3)
    best, best_t = cands[0], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        F.conv3d(x, w, padding=1)
        t0 = time.perf_counter()
        F.conv3d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best, avail


def cpu_train_step_factory(batch, seed=0):
    from oracle import genie_oracle as O
    import open_genie_b200 as og
    torch.manual_seed(seed)
    model = og.VideoTokenizer(og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, d_codebook=18, gan_loss_weight=0,
                              perc_loss_weight=0)          # CPU construction only: parameter shapes + default init
    sd = {k: v.detach().clone().contiguous().requires_grad_(v.dtype.is_floating_point)
          for k, v in model.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.AdamW(params)                         # the reference's default (genie/tokenizer.py:250)
    video = torch.randn(batch, 3, FRAMES, RES, RES)
    del model

    def step():
        loss, _, _, _ = O.tokenizer_forward(sd, og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, video, 18)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())
    return step


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads, avail = pick_cpu_threads()
    batch = args.cpu_batch
    step = cpu_train_step_factory(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    fps = batch * FRAMES / dt
    line = {
        'impl': 'reference', 'metric': 'videotokenizer_train_step_frames_per_sec', 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'MAGVIT2 VideoTokenizer train step (fwd+bwd+AdamW), {FRAMES}x{RES}x{RES} video, '
                               f'CPU sample of {batch} clip(s) per step'},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{batch} clip(s) x {FRAMES} frames per step, fp32, torch CPU ({threads} threads, '
                                   f'fastest of the {avail} available); '
                                   'oracle/genie_oracle.py restatement of the reference (pinned to reference outputs)'},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='clips per GPU (BASELINE configs[1]: 8)')
    ap.add_argument('--cpu-batch', type=int, default=1, help='clips per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying the captured step')
    args = ap.parse_args()
    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    import open_genie_b200 as og
    from open_genie_b200 import _lib, ops
    from open_genie_b200.ddp import GradBucketAllReducer

    torch.manual_seed(0)
    model = og.VideoTokenizer(og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, d_codebook=18, gan_loss_weight=0,
                              perc_loss_weight=0).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    opt = model.configure_optimizers()                      # FusedAdamW, AdamW defaults
    reducer = GradBucketAllReducer(model.parameters()) if world > 1 else None
    og.enable_zero_arena(True)   # every step below ends with zero_grad(set_to_none=True): the arena contract holds
    B = args.batch
    torch.manual_seed(1234 + rank)
    host_video = torch.randn(B, 3, FRAMES, RES, RES).pin_memory()
    dev_video = host_video.to(dev)
    h2d_bytes = host_video.numel() * 4

    def eager_step(video):
        loss = model.training_step(video, 0)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up (>= 3 steps) + capture of the whole step into one CUDA graph ----------------
    use_graph = (not args.no_graph) and world == 1   # N>1: eager launches (capturing the NCCL buckets hung in testing)
    for _ in range(max(args.warmup, 3)):
        eager_step(dev_video)
    barrier()
    if use_graph:
        from open_genie_b200.graph import GraphedTrainStep
        train_step = GraphedTrainStep(model, opt, dev_video, warmup=1, reducer=reducer)
        for _ in range(2):
            train_step(dev_video)
    else:
        train_step = eager_step
    barrier()

    # ---------------- device-resident leg (value) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        train_step(dev_video)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end-to-end leg: host batch in, loss out, every step ----------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        if use_graph:
            loss = train_step(host_video)                  # H2D from pinned memory into the static input + replay
        else:
            loss = train_step(host_video.to(dev, non_blocking=True))
        _ = loss.item()                                    # D2H read of the step's result
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))

    # ---------------- per-kernel timing pass (eager, CUDA events around every tensor-core launch) ----------
    # A replayed graph cannot carry timing events, so the roofline numbers come from the same step launched
    # eagerly right after the timed region (same process, same inputs, same kernels and grid sizes).
    if use_graph:
        model.zero_grad(set_to_none=True)
    prof_steps = min(args.steps, 3)
    ops.PROFILE = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eager_launch0 = _lib.launch_count()
    p0.record()
    for _ in range(prof_steps):
        eager_step(dev_video)
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    launches_per_step = (_lib.launch_count() - eager_launch0) / prof_steps
    prof, ops.PROFILE = ops.PROFILE, None
    if use_graph:
        launches = int(round(launches_per_step * args.steps))   # kernels executed by the replays of the timed region

    if rank == 0:
        peaks, peak_src = load_peaks()
        ms_step = ms_total / args.steps
        fps = world * B * FRAMES / (ms_step * 1e-3)
        fps_e2e = world * B * FRAMES / (ms_e2e / args.steps * 1e-3)
        # per-kernel roofline from the CUDA events recorded around every tensor-core launch
        kinds = {}
        for kind, flops, a, b, _shape in prof:
            k = 'og_conv_wgrad_kernel' if kind == 'wgrad' else 'og_conv_igemm_kernel'
            d = kinds.setdefault(k, {'ms': 0.0, 'flop': 0.0, 'launches': 0})
            d['ms'] += a.elapsed_time(b)
            d['flop'] += flops
            d['launches'] += 1
        kern = {}
        for k, d in kinds.items():
            kern[k] = {'launches_per_step': d['launches'] / prof_steps, 'ms_per_step': d['ms'] / prof_steps,
                       'tflops': d['flop'] / max(d['ms'], 1e-9) * 1e-9,
                       'share_of_step': (d['ms'] / prof_steps) / max(ms_step, 1e-9)}
        dom = max(kern, key=lambda k: kern[k]['ms_per_step']) if kern else None
        peak_tf = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
        roofline = None
        if dom:
            traffic, traffic_note = None, None
            try:    # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')))[dom]
                traffic = tj['traffic_bytes']
                traffic_note = (f"ncu --set full, one launch of {tj['shape']}: dram read {tj['dram_read_bytes']} + "
                                f"write {tj['dram_write_bytes']} B; algorithmic {tj['algorithmic_bytes']} B")
            except Exception:
                pass
            roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': kern[dom]['tflops'], 'peak': peak_tf,
                        'unit': 'TFLOP/s', 'frac': kern[dom]['tflops'] / peak_tf, 'traffic': traffic,
                        'traffic_note': traffic_note,
                        'peak_source': peak_src + ' sustained bf16', 'kernels': kern,
                        'conv_flop_per_step': sum(d['flop'] for d in kinds.values()) / prof_steps,
                        'timing': f'CUDA events around each launch over {prof_steps} eagerly launched steps after the '
                                  f'timed region ({ms_prof / prof_steps:.1f} ms/step eager)'}
        line = {
            'metric': 'videotokenizer_train_step_frames_per_sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[1]: MAGVIT2_ENC/DEC VideoTokenizer training step '
                                   '(fwd + bwd + AdamW, bf16 compute / fp32 master weights), d_codebook=18, '
                                   'GAN+perceptual terms off (reference runs offline only that way)',
                       'batch_per_gpu': B, 'global_batch': B * world, 'frames': FRAMES, 'resolution': RES,
                       'params': n_params, 'parallelism': f'dp{world}',
                       'launch': 'whole step replayed as one CUDA graph' if use_graph else 'eager launches',
                       'l2': 'no flush: every step streams several GB of activations (>> 126 MB L2)'},
            'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d_bytes * world,
                    'd2h_bytes_per_step': 4 * world},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads, avail = pick_cpu_threads()
            step = cpu_train_step_factory(args.cpu_batch)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {
                'value': args.cpu_batch * FRAMES / dt, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                'sample': f'one training step on {args.cpu_batch} clip(s) ({dt:.1f} s), fp32 torch CPU, '
                          f'{threads} threads (fastest of {avail} available), oracle port of the reference'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
