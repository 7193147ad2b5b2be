#!/usr/bin/env python
"""Benchmark of the GAViKO hot path (BASELINE.json metric: train / infer volumes/s, ViT-B/16 GAViKO, 1x120x160x160 volumes).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host CPU cores

A step = forward + focal loss + frozen-backbone backward + (N>1: NCCL all-reduce of the flat trainable gradient) + global-norm
clip + Adam, on one batch of synthetic volumes per GPU (weak scaling).  Prints ONE JSON line (rank 0).
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

GAVIKO_KW = dict(image_size=160, image_patch_size=16, frames=120, frame_patch_size=12, num_classes=5, channels=1, freeze_vit=True,
                 pool='cls', num_prompts=32, prompt_latent_dim=20, local_dim=20, local_k=[6, 6, 6], DHW=[10, 10, 10], dropout=0.1,
                 emb_dropout=0.1, attn_drop=0.2, proj_drop=0.2, share_factor=1, fp16=False)      # src/configs/gaviko.yaml:13-39
# algorithmic FLOPs per volume: forward, frozen-backbone backward
FLOPS = {'vit-t16': (23.5e9, 33.7e9), 'vit-b16': (222.4e9, 260.0e9), 'vit-l16': (742.2e9, 847.9e9)}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d, 'measured'
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms by ONE background `nvidia-smi -lms` process while the timed region runs
    (B200_PROFILING.md recipe); the median SM clock is taken over samples drawing > 300 W (= under load)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index), '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(',')]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
            self._t.join(timeout=2)

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        loaded = [r for r in self.rows if (num(r[2]) or 0) > 300.0] or self.rows
        sm = sorted(int(num(r[0])) for r in loaded if num(r[0]) is not None)
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        pw = [num(r[2]) for r in self.rows if num(r[2]) is not None]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=int(num(self.rows[0][1])) if self.rows and num(self.rows[0][1]) else None,
                    samples=len(self.rows), samples_under_load=len(loaded), power_w_max=max(pw) if pw else None, reasons=sorted(reasons))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(backbone, batch, steps, warmup, mode):
    """Times the reference algorithm (oracle restatement, fp32, torch CPU ops = what the reference itself dispatches to on CPU)
    for `batch` synthetic volumes: forward + focal loss + backward over the trainable set + clip_grad_norm_ + Adam (train.py:305-319; dropout off).
    Returns (seconds/step, threads)."""
    from oracle import gaviko_oracle as O
    from oracle.golden_fill import golden_fill, golden_labels, golden_volume
    from tests_support import sd_for_backbone
    torch.set_num_threads(os.cpu_count())
    sd, trainable = sd_for_backbone(backbone)
    golden_fill(sd, seed=0)
    for n in trainable:
        sd[n].requires_grad_(True)
    img = golden_volume(batch, 120, 160, 160)
    y = golden_labels(batch)
    params = [sd[n] for n in trainable]
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-8) if mode == 'train' else None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if mode == 'train':
            opt.zero_grad(set_to_none=True)
            logits = O.gaviko_forward(sd, img, backbone=backbone, num_prompts=32, frame_patch_size=12, image_patch_size=16, local_k=[6, 6, 6], DHW=[10, 10, 10])
            O.focal_loss(logits, y).backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        else:
            with torch.no_grad():
                O.gaviko_forward(sd, img, backbone=backbone, num_prompts=32, frame_patch_size=12, image_patch_size=16, local_k=[6, 6, 6], DHW=[10, 10, 10])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def gpu_eager_step_time(backbone, batch, steps, warmup, mode, dtype, dev):
    """The reference's algorithm in eager PyTorch ON THE SAME B200 (SURVEY 8d "beat it on the same box"): the oracle port moved to `dev` unchanged, so
    every op dispatches to ATen / cuBLAS as the reference's own modules would; fp32 as the scripts ship it, or `.to(bfloat16)`.  Train step =
    fwd + focal loss + bwd + clip_grad_norm_ + Adam over the trainable set (train.py:305-319), dropout off.  The oracle is the checker and sits outside
    our timed region; this leg is a reported baseline.  Returns seconds / step."""
    from oracle import gaviko_oracle as O
    from oracle.golden_fill import golden_fill, golden_labels, golden_volume
    from tests_support import sd_for_backbone
    sd, trainable = sd_for_backbone(backbone)
    golden_fill(sd, seed=0)
    sd = {k: v.to(dev, dtype) for k, v in sd.items()}
    params = []
    for n in trainable:
        sd[n].requires_grad_(True)
        params.append(sd[n])
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-8) if mode == 'train' else None
    img = golden_volume(batch, 120, 160, 160).to(dev, dtype)
    y = golden_labels(batch).to(dev)
    kw = dict(backbone=backbone, num_prompts=32, frame_patch_size=12, image_patch_size=16, local_k=[6, 6, 6], DHW=[10, 10, 10])

    def step():
        if mode == 'train':
            opt.zero_grad(set_to_none=True)
            O.focal_loss(O.gaviko_forward(sd, img, **kw).float(), y).backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        else:
            with torch.no_grad():
                O.gaviko_forward(sd, img, **kw)
    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 2
    sec, threads = cpu_reference_step_time(args.backbone, batch, args.steps, args.warmup, args.mode)
    value = batch / sec
    line = dict(impl='reference', metric=metric_name(args), value=value, unit='volumes/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=workload(args, batch, 1, reference=True), gpu_launches=0,
                cpu_baseline=dict(value=value, unit='volumes/s', cores=threads, kind='port',
                                  sample=f'{args.steps} steps of batch {batch} (fwd + focal loss + bwd + clip + Adam, dropout off) of the same {args.backbone} GAViKO workload, fp32, torch CPU'),
                e2e=dict(value=value, unit='volumes/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def metric_name(args):
    return f"{'train' if args.mode == 'train' else 'infer'} volumes/sec {args.backbone} GAViKO"


def workload(args, batch, world, reference=False):
    if args.mode != 'train':
        what = 'batched inference forward'
    elif reference:       # the CPU arm times what the reference's loop does per step (train.py:305-319), dropout off
        what = 'training step (fwd + focal loss + frozen-backbone bwd + clip + Adam; dropout off)'
    else:
        what = 'training step (fwd + focal loss + frozen-backbone bwd + clip + Adam)'
    return dict(workload=f"GAViKO {args.backbone} {what}, "
                         'synthetic 1x120x160x160 volumes, random-init weights', backbone=args.backbone, per_gpu_batch=batch, global_batch=batch * world,
                volume='1x120x160x160 fp32', num_prompts=32, tokens=1033, parallelism=f'dp{world}',
                l2_policy=f'inputs larger than L2: {batch * 12.288:.0f} MB of volumes + >1 GB of activations per step vs 126 MB L2')


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from gaviko_b200 import _lib as L
    from gaviko_b200 import ops
    from gaviko_b200.losses.focal_loss import FocalLoss
    from gaviko_b200.model.gaviko import Gaviko
    from gaviko_b200.optim import FlatAdam

    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (gaviko_b200 has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    import contextlib
    import io
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = Gaviko(**GAVIKO_KW, backbone=args.backbone, compute_dtype=args.dtype)
    model = model.to(dev)
    crit = FocalLoss(gamma=1.2)
    train = args.mode == 'train'
    if train:
        model.train()
        opt = FlatAdam(model.parameters(), lr=1e-4, eps=1e-8, max_grad_norm=1.0, world_size=world, model=model)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-4, total_steps=2 * (args.steps + args.warmup) + 16, pct_start=0.3, div_factor=10.0,
                                                    final_div_factor=1000.0, anneal_strategy='cos', three_phase=False)
    else:
        model.eval()
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    # two pinned host batches (double buffer) for the end-to-end leg; one resident device batch for the kernel-only leg
    host = [torch.rand(B, 1, 120, 160, 160, generator=g).pin_memory() for _ in range(2)]
    labels_host = torch.randint(0, 5, (B,), generator=g).pin_memory()
    x_dev = host[0].to(dev)
    y_dev = labels_host.to(dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    gstep = None
    launches_per_replay = 0
    launch_mode = 'eager (one launch per kernel)'
    if train and not args.eager:
        # default: the whole step (forward + loss + backward + all-reduce + clip + Adam) captured once as ONE CUDA graph and replayed
        from gaviko_b200.graph import GraphedTrainStep
        n0 = L.launch_count()
        try:
            gstep = GraphedTrainStep(model, crit, opt, x_dev, y_dev, warmup=2)
            launches_per_replay = (L.launch_count() - n0) // 3          # 2 warm-up steps + the captured one enqueue the same kernels
            launch_mode = 'one CUDA graph per step'
        except Exception as ex:  # noqa: BLE001   (reported in config.launch: the same kernels then run launch by launch)
            gstep = None
            launch_mode = f'eager (graph capture failed: {type(ex).__name__}: {ex})'[:200]

    gfwd = None
    if not train and not args.eager:
        from gaviko_b200.graph import GraphedForward
        n0 = L.launch_count()
        try:
            gfwd = GraphedForward(model, x_dev, warmup=2)
            launches_per_replay = (L.launch_count() - n0) // 3
            launch_mode = 'one CUDA graph per forward'
        except Exception as ex:  # noqa: BLE001
            gfwd = None
            launch_mode = f'eager (graph capture failed: {type(ex).__name__}: {ex})'[:200]

    def step(x, y):
        if gfwd is not None:
            return gfwd(x)
        if gstep is not None:
            loss = gstep(x, y)
            sched.step()
            return loss
        if train:
            logits = model(x)
            loss = crit(logits, y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            sched.step()
            return loss
        with torch.no_grad():
            return model(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM -------------------------------------------------------
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    barrier()
    gemm_events = []
    ops.GEMM_HOOK = gemm_events if rank == 0 else None
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0 + launches_per_replay * args.steps      # kernel nodes of a graph replay are not seen by the library's counter
    ops.GEMM_HOOK = None
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()

    # ---- leg 2: end to end through the public API with host buffers ---------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(x_dev) for _ in range(2)]
    ybufs = [torch.empty_like(y_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs[i % 2].copy_(host[i % 2], non_blocking=True)
            ybufs[i % 2].copy_(labels_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            out = step(bufs[i % 2], ybufs[i % 2])
            consumed[i % 2].record()
            res = out if train else out.argmax(1).float().sum()
            loss_host.copy_(res.detach().float().reshape(()), non_blocking=True)   # D2H read of the step's result
        torch.cuda.current_stream().synchronize()

    for ev in consumed:
        ev.record()
    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()

    if rank == 0:
        pk, pk_src = peaks()
        fwd_f, bwd_f = FLOPS[args.backbone]
        flops_per_vol = fwd_f + (bwd_f if train else 0.0)
        value = world * B * args.steps / (ms / 1e3)
        e2e_value = world * B * args.steps / (ms_e2e / 1e3)
        # roofline of the dominant kernel class: the tcgen05 GEMM (all launches inside the timed region, CUDA events on the launch stream)
        if gstep is not None or gfwd is not None:      # no per-kernel events inside a graph replay: time the same GEMM launches over two eager steps after the timed region
            ops.GEMM_HOOK = gemm_events
            for _ in range(2):
                if train:
                    loss = crit(model(x_dev), y_dev); opt.zero_grad(); loss.backward()
                else:
                    with torch.no_grad():
                        model(x_dev)
            torch.cuda.synchronize()
            ops.GEMM_HOOK = None
        gsec = sum(a.elapsed_time(b) for a, b, _ in gemm_events) / 1e3
        gflop = sum(f for _, _, f in gemm_events)
        peak = pk['bf16_tflops_sustained'] if args.dtype == 'bf16' else 80.0
        traffic = None      # DRAM bytes per GEMM launch from the committed ncu capture of this configuration (profiles/), else null
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'ncu_r02_gemm_traffic.json')))
            if (tj['batch'], tj['backbone'], tj['dtype']) == (B, args.backbone, args.dtype) and train:
                traffic = tj['dram_bytes_per_launch']
        except Exception:  # noqa: BLE001
            pass
        roof = dict(bound='tensor', kernel='gemm_bf16_sm100_kernel' if args.dtype == 'bf16' else 'gemm_f32_kernel', achieved=(gflop / gsec / 1e12) if gsec > 0 else None,
                    peak=peak, unit='TFLOP/s', frac=(gflop / gsec / 1e12 / peak) if gsec > 0 else None, traffic=traffic, peak_source=f'{pk_src} (bf16_tflops_sustained)',
                    launches=len(gemm_events), share_of_step=(gsec * 1e3 / (2 if (gstep is not None or gfwd is not None) else args.steps)) / (ms / args.steps) if ms > 0 else None,
                    step_tensor_frac=world and (B * args.steps * flops_per_vol / (ms / 1e3) / 1e12 / peak))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, threads = cpu_reference_step_time(args.backbone, 2, 8, 1, args.mode)
                cpu = dict(value=2 / sec, unit='volumes/s', cores=threads, kind='port',
                           sample=f'1 warm-up + 8 timed steps of batch 2 (~10 s) of the same {args.backbone} GAViKO {args.mode} workload (oracle port, fp32, torch CPU threads)')
            except Exception as ex:  # noqa: BLE001
                cpu = dict(value=None, unit='volumes/s', cores=os.cpu_count(), kind='port', sample=f'failed: {ex}')
        eager = None
        if world == 1 and not args.no_gpu_eager_baseline:
            del model, x_dev, bufs
            if train:
                del opt, sched
            torch.cuda.empty_cache()
            eager = {}
            eb = min(B, 8)         # the eager path materialises (B, 12, T, T) scores per layer and keeps them for backward: ~1.2 GB / volume fp32
            for name, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
                try:
                    sec = gpu_eager_step_time(args.backbone, eb, 3, 2, args.mode, dt, dev)
                    eager[name] = dict(value=eb / sec, unit='volumes/s', ms_per_step=sec * 1e3,
                                       sample=f'2 warm-up + 3 timed steps of batch {eb}: oracle port of the reference in eager PyTorch on this GPU ({name}; ATen / cuBLAS; dropout off)')
                except Exception as ex:  # noqa: BLE001
                    eager[name] = dict(value=None, unit='volumes/s', sample=f'failed: {type(ex).__name__}: {ex}'[:300])
                torch.cuda.empty_cache()
        line = dict(metric=metric_name(args), value=value, unit='volumes/s', n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype=args.dtype, data='synthetic', config=dict(workload(args, B, world), launch=launch_mode), clocks=clocks,
                    e2e=dict(value=e2e_value, unit='volumes/s', h2d_bytes_per_step=world * (B * 120 * 160 * 160 * 4 + B * 8), d2h_bytes_per_step=world * 4,
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=launches, roofline=roof, cpu_baseline=cpu, gpu_eager_baseline=eager,
                    tflops_algorithmic=world * B * args.steps * flops_per_vol / (ms / 1e3) / 1e12)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--backbone', default='vit-b16')
    ap.add_argument('--batch', type=int, default=64, help='volumes per GPU per step (SURVEY 8d: 16 | 32 | 64)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--graph', action='store_true', help='(default for training) replay the whole step as ONE CUDA graph (gaviko_b200.graph.GraphedTrainStep)')
    ap.add_argument('--eager', action='store_true', help='training: launch kernel by kernel from Python instead of replaying the captured CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-eager-baseline', action='store_true', help='skip timing the eager-PyTorch port of the reference on this GPU (N=1 only)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
