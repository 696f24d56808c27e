#!/usr/bin/env python
"""Benchmark of the D2C-VAE continuous-decoding path (BASELINE.json metric).

Default workload = BASELINE.json configs[1]: arbitrary-resolution image decode of
AFHQ-shape latents (3 PE planes 64^2/128^2/256^2 x 64 ch) at 1024x1024 and
2048x2048 query grids, batch 64 per GPU.  One "step" = one pass of the hot path over
that batch: one decode at 1024^2 then one at 2048^2 (3.36e8 coordinates / GPU).

Prints ONE JSON line (rank 0).  `value` = decoded coordinates / s, whole job, inputs
resident in HBM; `e2e` = same through the public module API with pinned HOST buffers
(planes H2D + RGB D2H inside the timed region).  `--impl reference` times the CPU
restatement of the reference decoder (oracle/, the reference itself is Python and
cannot travel to the GPU box) on a bounded sample, all host threads.

`--workload c1 | video | occupancy | nerf` prints the same line (value, e2e, clocks, roofline,
cpu_baseline, gpu_eager_baseline) for the other BASELINE configs (configs[0], [2], [3], [4]) and
`--workload mesh` the occupancy post-step (128^3 grid -> mesh); the default (headline) line also carries
`other_configs`: the device-resident throughput + roofline fraction of those four configs, measured in the
same command;
`gpu_eager_baseline` = the same oracle restatement as eager PyTorch fp32 (TF32 off) on the SAME
B200 in the same run: the like-for-like GPU comparator.  With N > 1 GPUs the
image line also carries `strong_scaling`: configs[1] with the batch of 64 split over the ranks by
ddmi_b200.sharding, timed without and with the output all-gather.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_COORD = {'image': 1908224, 'video': 1713664, 'occupancy': 1255424, 'nerf': 1104384}
METRIC = "decoded INR coords/sec"
UNIT = "coords/s"


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1415.7), d.get('hbm_gbs', 6452.8), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.FIELDS}',
                                       '--format=csv,noheader,nounits', '-lms', '200'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
            # nvidia-smi's start-up stalls the device for ~0.2 s: let it print its first sample before anything is timed
            t0 = time.time()
            while time.time() - t0 < 2.0 and os.path.getsize(self.f.name) == 0:
                time.sleep(0.05)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                pw = float(r[2])
                if pw > 300:      # under load
                    sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.strip().lower() == 'active':
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples_under_load": len(sm)}


def build_mlp():
    """Random-init weights of the AFHQ image decoder (afhq.yaml:44-48) under the reference's default seed."""
    import ddmi_b200
    torch.manual_seed(777)
    m = ddmi_b200.MLP(in_ch=2, latent_dim=64, out_ch=3, ch=256)
    g = torch.Generator().manual_seed(778)
    for name, p in m.named_parameters():
        if name.endswith('activate.bias') or name == 'torgb.bias':
            p.data = 0.1 * torch.randn(p.shape, generator=g)
    return m.eval()


def state_dict32(m):
    return {k: v.detach().float().cpu() for k, v in m.state_dict().items()}


def image_workload(batch, device, pinned=False):
    from ddmi_b200 import convert_to_coord_format_2d, get_scale_injection
    g = torch.Generator().manual_seed(777)
    planes = [torch.randn(batch, 64, s, s, generator=g) for s in (64, 128, 256)]
    grids = []
    for R in ARGS.res:
        e = (R - 1) / R
        grids.append((convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(device),
                      get_scale_injection(R), R))
    if pinned:
        planes = [p.pin_memory() for p in planes]
    return planes, grids


# per precision: what the tensor pipe executes per algorithmic flop, and how long that takes in units of one bf16 pass
PREC_INFO = {
    'bf16x3': {'dtype': "bf16x3 (bf16 hi/lo split operands, 3 MMAs per product, fp32 accumulate)",
               'kernel': "image_umma_kernel<pair, bf16x3>", 'executed': 3.0, 'tensor_time': 3.0,
               'note': "the bf16x3 split executes 3x that on the tensor pipe"},
    'f16f8': {'dtype': "f16f8 (fp16 main term + two e4m3 correction terms at the FP8 rate, fp32 accumulate)",
              'kernel': "image_umma_kernel<pair, f16f8>", 'executed': 3.0, 'tensor_time': 2.0,
              'note': "the f16f8 split executes 1x that in fp16 + 2x in FP8 (= 2 bf16-pass times) on the tensor pipe"},
    'fp32': {'dtype': "f32", 'kernel': "fp32::image_kernel", 'executed': 1.0, 'tensor_time': 0.0,
             'note': "CUDA-core fp32 FMAs"},
}


def eager_sync_time(fn, repeats=2):
    """best-of wall time of fn() on the current CUDA device (synchronised)."""
    best = float('inf')
    for _ in range(repeats + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def gpu_eager_image(dev, res=1024, batch=2):
    """The oracle restatement of MLP.forward as eager PyTorch on the GPU (fp32, TF32 off): the like-for-like GPU baseline."""
    from oracle import ddmi_oracle as orc   # the checker, timed here as a baseline only
    from ddmi_b200 import convert_to_coord_format_2d, get_scale_injection
    tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = {k: v.to(dev) for k, v in state_dict32(build_mlp()).items()}
        g = torch.Generator().manual_seed(777)
        planes = [torch.randn(1, 64, s, s, generator=g).to(dev) for s in (64, 128, 256)]
        e = (res - 1) / res
        coords = convert_to_coord_format_2d(1, res, res, hstart=-e, hend=e, wstart=-e, wend=e).to(dev)
        si = get_scale_injection(res)

        def run():
            for _ in range(batch):                      # one item at a time: (1, 322, res, res) fp32 activations = 1.35 GB
                orc.image_decode(sd, coords, planes, si)
        dt = eager_sync_time(run)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
    return {"value": batch * res * res / dt, "unit": UNIT, "kind": "oracle restatement of the reference decoder, eager PyTorch fp32 "
            "(TF32 off) on this GPU", "sample": f"{batch} items @ {res}x{res}, one item per call, best of 3"}


def pack_timing(mlp, grids, dev):
    """Host cost of a NEW scale injection value (arbitrary-resolution decode changes si per resolution): fold + pack on the CPU
    + three H2D copies; cached per (precision, si) afterwards."""
    out = {}
    for c, si, R in grids:
        mlp.invalidate_packed()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mlp(c[:, :, :1, :128], hdbf=[torch.zeros(1, 64, 4, 4, device=dev)] * 3, si=si)
        torch.cuda.synchronize()
        out[str(R)] = (time.perf_counter() - t0) * 1e3
    return out


def strong_scaling(args, mlp, dev, world, rank, grids):
    """configs[1] with its batch of 64 SPLIT over the ranks (ddmi_b200.sharding): time to decode, and to decode + assemble the
    full signal on every rank with one all_gather_into_tensor of the owned slabs (NCCL)."""
    import torch.distributed as dist
    from ddmi_b200 import sharding
    B = 64
    g = torch.Generator().manual_seed(777)
    planes = [torch.randn(B, 64, s, s, generator=g).to(dev) for s in (64, 128, 256)]
    res = {}
    for c, si, R in grids:
        if R > 1024 and world < 4:
            continue                                           # the assembled 2048^2 signal is 3.2 GB per rank: keep it to N >= 4
        row = {}
        plan = sharding.plan_units(B, R, world)

        def timed(fn, reps):
            best = float('inf')
            out = None
            for _ in range(reps + 1):                          # the first repetition warms up (allocations after empty_cache)
                dist.barrier()
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                out = fn()
                t1.record()
                torch.cuda.synchronize()
                t = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = min(best, float(t[0]))
                row.setdefault('all_ms', []).append(round(float(t[0]), 2))
            return best, out

        row['decode_ms'], mine = timed(lambda: sharding.decode_image_sharded(mlp, c, planes, si=si, gather=False), 2)
        # the one collective, timed on its own: every rank contributes only its own items (all_gather_into_tensor + placement)
        row['gather_ms'], full = timed(lambda: sharding.assemble(mine, plan, (B, 3, R, R), 2, None, 'all'), 3)
        assert tuple(full.shape) == (B, 3, R, R)
        del full, mine
        row['decode_and_gather_ms'] = row['decode_ms'] + row['gather_ms']
        row['coords_per_s'] = B * R * R / (row['decode_ms'] * 1e-3)
        row['coords_per_s_with_gather'] = B * R * R / (row['decode_and_gather_ms'] * 1e-3)
        row['gathered_bytes_per_rank'] = B * 3 * R * R * 4 * (world - 1) // world
        row['gather_gb_per_s_per_rank'] = row['gathered_bytes_per_rank'] / (row['gather_ms'] * 1e-3) / 1e9
        res[str(R)] = row
    return {"scaling": "strong", "batch_total": B, "n_gpus": world, "per_grid": res,
            "note": "times = max over ranks, CUDA events, best of 2-3; gather = one all_gather_into_tensor of each rank's own items + placement into the (B,3,R,R) result"}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.set_grad_enabled(False)
    B = args.batch
    mlp = build_mlp().to(dev)
    mlp.precision = args.precision
    host_planes, grids = image_workload(B, dev, pinned=True)
    planes = [p.to(dev) for p in host_planes]
    coords_per_step = sum(B * R * R for _, _, R in grids)
    planes_gb = sum(p.numel() * 4 for p in planes) / 1e9

    def step():
        outs = []
        for c, si, R in grids:
            outs.append(mlp(c, hdbf=planes, si=si, store=args.store))
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None     # started BEFORE the warm-up: its start-up stalls the device briefly
    for _ in range(args.warmup):
        step()
    barrier()
    # per-launch device times of the dominant kernel, on the launching (current) stream
    evs = []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        for c, si, R in grids:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = mlp(c, hdbf=planes, si=si, store=args.store)
            e1.record()
            evs.append((e0, e1, B * R * R))
            del out
    t_end.record()
    barrier()
    ms = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None
    launch_ms = [(a.elapsed_time(b), n) for a, b, n in evs]

    # ---- end to end through the public API with host buffers ----
    if args.store == 'u8':
        out_host = [torch.empty((B, R, R, 3), dtype=torch.uint8).pin_memory() for _, _, R in grids]
    else:
        out_host = [torch.empty((B, 3, R, R), dtype=torch.float32).pin_memory() for _, _, R in grids]
    h2d = sum(p.numel() * 4 for p in host_planes)            # the step's latents are uploaded once
    d2h = sum(o.numel() * o.element_size() for o in out_host)
    copy_stream = torch.cuda.Stream(device=dev)
    up_stream = torch.cuda.Stream(device=dev)
    CH = max(1, min(args.e2e_chunk, B))                      # batch items per pipelined chunk

    def e2e_step():
        # one step = upload this step's latents (pinned -> HBM), decode them at every query grid through the
        # public module call, read every RGB grid back to pinned host memory.  The batch is walked in chunks of CH
        # items: the upload of chunk k+1 (side stream) and the read-back of chunk k (another side stream) overlap
        # the decode of their neighbours, so only the first upload and the last read-back are exposed.
        main = torch.cuda.current_stream()
        chunks = [(k, min(k + CH, B)) for k in range(0, B, CH)]
        ups = []
        with torch.cuda.stream(up_stream):
            for a, b in chunks:
                dp = [p[a:b].to(dev, non_blocking=True) for p in host_planes]
                ev = torch.cuda.Event()
                ev.record(up_stream)
                ups.append((dp, ev))
        for (a, b), (dp, ev) in zip(chunks, ups):
            main.wait_event(ev)
            for t in dp:
                t.record_stream(main)
            for i, (c, si, R) in enumerate(grids):
                o = mlp(c, hdbf=dp, si=si, store=args.store)
                copy_stream.wait_stream(main)
                with torch.cuda.stream(copy_stream):
                    out_host[i][a:b].copy_(o, non_blocking=True)
                o.record_stream(copy_stream)
        main.wait_stream(copy_stream)
        main.synchronize()

    e2e_step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 3))
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    strong = None
    if world > 1 and not args.no_strong:
        del planes
        torch.cuda.empty_cache()
        strong = strong_scaling(args, mlp, dev, world, rank, grids)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tf_peak, hbm_peak, which = peaks()
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, 'profiles', {'bf16x3': 'r01b_image_traffic.json', 'f16f8': 'r01e_image_traffic_f16f8.json'}
                      .get(args.precision, 'none'))
    if os.path.exists(tp):
        t = json.load(open(tp))
        traffic = t['dram_bytes_read'] + t['dram_bytes_write']
        traffic_note = (f"ncu dram bytes of ONE profiled launch ({t['launch']}): {traffic} B vs {t['algorithmic_bytes']} B "
                        f"algorithmic (planes once + 12 B/coord out); tensor pipe {t['tensor_pipe_active_pct']} % active")
    value = coords_per_step * world * args.steps / (ms * 1e-3)
    tot_flop = sum(n * FLOP_PER_COORD['image'] for _, n in launch_ms)
    tot_s = sum(t for t, _ in launch_ms) * 1e-3
    achieved = tot_flop / tot_s / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": PREC_INFO[args.precision]['dtype'],
        "data": "synthetic",
        "config": {"workload": f"AFHQ-shape image D2C-VAE decode (BASELINE configs[1]): planes 64^2/128^2/256^2 x64ch, "
                               f"query grids {'+'.join(f'{R}x{R}' for R in args.res)}, batch {B} per GPU",
                   "coords_per_step_per_gpu": coords_per_step, "precision": args.precision,
                   "store": args.store + {"f32": " (the reference's output: (B,3,h,w) fp32)", "clamp": " (clamp(-1,1) fused)",
                                           "u8": " (uint8 (B,h,w,3) = trunc((clamp(x,-1,1)+1)*127.5) fused: the callers' epilogue)"}[args.store],
                   "l2": "inputs larger than L2 (planes %.2f GB per GPU; no flush)" % planes_gb,
                   "sharding": "batch items per rank, no collective"},
        "e2e": {"value": coords_per_step * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "pipeline": f"batch walked in chunks of {CH} items; uploads / read-backs on side streams"},
        "gpu_launches": args.steps * len(grids),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": achieved / tf_peak, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": which + " (bf16 sustained)",
                     "kernel": PREC_INFO[args.precision]['kernel'],
                     "note": "achieved = ALGORITHMIC flops (1,908,224 / coord, as-written layers); "
                             + PREC_INFO[args.precision]['note'],
                     "executed_tflops": achieved * PREC_INFO[args.precision]['executed'],
                     "bf16_equivalent_tensor_time": PREC_INFO[args.precision]['tensor_time'],
                     "avg_launch_ms": {str(n): statistics.mean(t for t, m in launch_ms if m == n) for n in sorted({m for _, m in launch_ms})}},
    }
    if world == 1:
        # Self-check of the measured kernel (not timed): the same module on a sample of the workload (4 items @ 256x256) through
        # the exact-arithmetic fp32 CUDA-core kernels; and, for context, the other tensor-core operand scheme's throughput.
        c0, si0, _ = grids[0]
        cs = torch.nn.functional.interpolate(c0, size=(256, 256), mode='bilinear', align_corners=True)
        sample = [p[:4] for p in planes]
        mlp.precision = args.precision
        got = mlp(cs, hdbf=sample, si=si0)
        mlp.precision = 'fp32'
        exact = mlp(cs, hdbf=sample, si=si0)
        line["parity"] = {"max_abs_vs_fp32_kernel": float((got - exact).abs().max()), "tolerance": 1e-3,
                          "sample": "4 items @ 256x256, same planes and weights"}
        other = {'f16f8': 'bf16x3', 'bf16x3': 'f16f8'}.get(args.precision)
        if other and not args.no_cpu_baseline:        # context legs are skipped together (profiler runs, quick checks)
            mlp.precision = other
            step()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                step()
            a1.record()
            torch.cuda.synchronize()
            line["other_precision"] = {"precision": other, "value": coords_per_step * 2 / (a0.elapsed_time(a1) * 1e-3),
                                       "unit": UNIT, "steps": 2}
        mlp.precision = args.precision
    if world == 1 and not args.no_cpu_baseline:
        del planes
        torch.cuda.empty_cache()
        # the other BASELINE configs, device-resident, in this same command (their full lines: --workload c1 | video | ...)
        line["other_configs"] = {k: quick_other(k, args, dev) for k in ('c1', 'video', 'occupancy', 'nerf')}
        line["cpu_baseline"] = cpu_baseline(args.cpu_res, args.cpu_batch)
        line["gpu_eager_baseline"] = gpu_eager_image(dev)
        line["gpu_eager_baseline"]["speedup_of_value"] = value / line["gpu_eager_baseline"]["value"]
        line["pack_ms_per_new_si"] = pack_timing(mlp, grids, dev)
    if world > 1:
        line["strong_scaling"] = strong
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class OtherWorkload:
    """One of the non-headline BASELINE configs: module, pinned host planes, fixed device-side query set, a decode over a
    range of items, and the oracle on a bounded sample (CPU and eager-GPU baselines)."""

    def __init__(self, kind, args, dev):
        import numpy as np
        import ddmi_b200
        from ddmi_b200 import nerf_helpers as nh
        self.kind, self.dev, self.precision = kind, dev, args.precision
        g = torch.Generator().manual_seed(777)
        torch.manual_seed(777)
        pin = lambda t: t.pin_memory()
        if kind == 'c1':
            self.flop_kind, self.kernel = 'image', 'image_umma_kernel'
            B = self.B = args.batch if args.batch != 64 else 4
            self.m = build_mlp().to(dev)
            self.host = [pin(torch.randn(B, 64, s, s, generator=g)) for s in (64, 128, 256)]
            R, e = 256, 255 / 256
            self.coords = ddmi_b200.convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(dev)
            self.n_per_item = R * R
            self.out_item = ((3, R, R), torch.float32)
            self.decode = lambda planes: self.m(self.coords, hdbf=planes, si=1.0)
            self.desc = f"AFHQ-shape image D2C-VAE decode (BASELINE configs[0]): planes 64^2/128^2/256^2 x64ch, 256x256 grid, si=1, batch {B}"
        elif kind == 'occupancy':
            self.flop_kind, self.kernel = 'occupancy', 'occupancy_umma_kernel'
            B = self.B = args.batch if args.batch != 64 else 32
            self.m = ddmi_b200.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256).to(dev)
            for blk in (self.m.net_res1, self.m.net_res2, self.m.net_res3, self.m.net_res4):
                torch.nn.init.kaiming_uniform_(blk.fc_1.weight, a=5 ** 0.5)
            self.host = [[pin(torch.randn(B, 64, s, s, generator=g)) for s in (16, 32, 64)] for _ in range(3)]
            # the dense grid is queried as a lattice (MLP3D.decode_logits_lattice = the mesh generator's query,
            # 1.1 * make_3d_grid((-.5,)*3, (.5,)*3, (128,)*3), logits bit-identical to the point list), the random points as a point list
            self.axis = (1.1 * torch.linspace(-0.5, 0.5, 128)).to(dev)
            self.pts = ((torch.rand(100000, 3, generator=g) - 0.5) * 1.1).to(dev)
            self.n_per_item = 128 ** 3 + self.pts.shape[0]
            self.out_item = ((self.n_per_item,), torch.float32)

            def decode(planes):
                b = planes[0][0].shape[0]
                grid = self.m.decode_logits_lattice((self.axis,) * 3, planes).reshape(b, -1)
                return torch.cat([grid, self.m.decode_logits(self.pts[None].expand(b, -1, -1), planes)], dim=1)
            self.decode = decode
            self.launches = 4              # lattice tables, lattice decode, point-list decode, the concatenation
            self.desc = (f"ShapeNet-shape occupancy decode (BASELINE configs[3]): triplanes 16^2/32^2/64^2 x64ch, 128^3 grid "
                         f"(lattice query) + 100k random points (point list), batch {B}")
        elif kind == 'video':
            self.flop_kind, self.kernel = 'video', 'video_umma_kernel'
            B = self.B = args.batch if args.batch != 64 else 16
            self.m = ddmi_b200.MLPVideo(in_ch=2, latent_dim=64, out_ch=3, ch=256).to(dev)
            for blk in (self.m.net_res1, self.m.net_res2, self.m.net_res3, self.m.net_res4):
                torch.nn.init.kaiming_uniform_(blk.fc_1.weight, a=5 ** 0.5)
            self.host = [[pin(torch.randn(B, 64, s, s, generator=g)) for s in (64, 128, 256)],
                         [pin(torch.randn(B, 64, 16, s, generator=g)) for s in (64, 128, 256)],
                         [pin(torch.randn(B, 64, 16, s, generator=g)) for s in (64, 128, 256)]]
            c = ddmi_b200.convert_to_coord_format_3d(1, 256, 256, 16, hstart=-255 / 256, hend=255 / 256, wstart=-255 / 256,
                                                     wend=255 / 256, tstart=-15 / 16, tend=15 / 16)
            self.coords = {k: v.to(dev) for k, v in c.items()}
            self.n_per_item = 256 * 256 * 16
            self.out_item = ((3, 16, 256, 256), torch.float32)
            self.decode = lambda planes: self.m(self.coords, planes)
            self.launches = 2              # feature tables + decode
            self.desc = f"SkyTimelapse-shape video decode (BASELINE configs[2]): xy/yt/xt planes, 256x256x16 volume, batch {B}"
        else:
            self.flop_kind, self.kernel = 'nerf', 'nerf_umma_kernel'
            B = self.B = args.batch if args.batch != 64 else 16
            self.m = ddmi_b200.MLPNeRF(D=6, W=256, in_channels_xyz=159, skips=[2, 4], in_channels_dir=27).to(dev)
            self.host = [pin(torch.randn(B, 32, 64, 64, generator=g)) for _ in range(3)]
            H = W = 128
            focal = .5 * W / np.tan(.5 * 0.6911112070083618)
            K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
            ro, rd = nh.get_rays(H, W, K, nh.pose_spherical(40.0, -20, 5)[:3, :4], dev)
            vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
            self.rays = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), 2. * torch.ones(H * W, 1, device=dev),
                                   6. * torch.ones(H * W, 1, device=dev), vd], -1)
            self.n_per_item = H * W * 128
            self.out_item = ((H * W, 3), torch.float32)
            self.decode = lambda planes: nh.render_rays_fused(self.rays, dict(zip(('xy', 'yz', 'xz'), planes)), self.m, 128, True,
                                                              precision=args.precision)
            self.desc = (f"srn-cars-shape NeRF decode (BASELINE configs[4]): triplane 64^2 x32ch, 128x128 rays x 128 samples, "
                         f"composited, batch {B} objects (coordinates = ray samples)")
        self.m.precision = args.precision

    def slice(self, host, a, b, dev=None):
        f = (lambda t: t[a:b].to(dev, non_blocking=True)) if dev is not None else (lambda t: t[a:b])
        return [([f(t) for t in x] if isinstance(x, list) else f(x)) for x in host]

    def h2d_bytes(self):
        return sum(t.numel() * 4 for x in self.host for t in (x if isinstance(x, list) else [x]))

    def oracle_sample(self, dev, n_target):
        """The oracle restatement on `dev` over a bounded sample of this workload -> (coords decoded, callable)."""
        from oracle import ddmi_oracle as orc   # the checker, timed here as a baseline only
        sd = {k: v.detach().float().to(dev) for k, v in self.m.state_dict().items()}
        one = self.slice(self.host, 0, 1)
        mv = lambda x: [mv(t) for t in x] if isinstance(x, list) else x.to(dev)
        planes = mv(one)
        if self.kind == 'c1':
            c = self.coords.to(dev)
            return self.n_per_item, lambda: orc.image_decode(sd, c, planes, 1.0), "1 item @ 256x256"
        if self.kind == 'occupancy':
            n = min(n_target, 100000)
            p = self.pts[-n:][None].to(dev)                    # the reference's own chunk size (generation.py:130-144)
            return n, lambda: orc.occupancy_logits(sd, p, planes), f"1 item, {n} random points (one eval_points chunk)"
        if self.kind == 'video':
            rows = max(1, min(256, n_target // (16 * 256)))
            c = {k: v.to(dev) for k, v in self.coords.items()}
            sub = {'xy': c['xy'][:, :, :rows], 'yt': c['yt'][:, :, :, :rows], 'xt': c['xt']}
            return 16 * rows * 256, (lambda: orc.video_decode(sd, sub, planes, thw=(16, rows, 256))), \
                f"1 item, 16 frames x {rows} rows x 256"
        nr = max(1, min(self.rays.shape[0], n_target // 128))
        r = self.rays[:nr].to(dev)
        fea = dict(zip(('xy', 'yz', 'xz'), planes))
        return nr * 128, lambda: orc.nerf_render_rays(sd, r, fea, 128, True), f"1 object, {nr} rays x 128 samples"


def quick_other(kind, args, dev):
    """Device-resident throughput of one of the non-headline BASELINE configs (3 timed steps after 2 warm-ups), for the
    `other_configs` key of the headline line: every config is then measured in the same driver-run command."""
    a = argparse.Namespace(**vars(args))
    a.batch = 64                                   # = "the config's own batch" in OtherWorkload
    W = OtherWorkload(kind, a, dev)
    planes = W.slice(W.host, 0, W.B, dev)
    for _ in range(2):
        W.decode(planes)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        W.decode(planes)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    coords = W.B * W.n_per_item
    tf_peak, _, _ = peaks()
    value = coords / (ms * 1e-3)
    out = {"workload": W.desc, "value": value, "unit": UNIT, "ms_per_step": ms, "kernel": W.kernel,
           "roofline_frac": value * FLOP_PER_COORD[W.flop_kind] / 1e12 / tf_peak}
    del planes, W
    torch.cuda.empty_cache()
    return out


def run_other(args):
    """configs[0], [2], [3], [4]: the same line as the headline (value, e2e, clocks, roofline, cpu / eager-GPU baselines).
    Under torchrun (N ranks, one per GPU): weak scaling like the headline -- every rank decodes the config's batch of its own
    items (the path shards by item with no exchange), barrier + synchronize around the timed region, time = max over ranks,
    value = N x the per-rank coordinates / that time; rank 0 prints."""
    import torch.distributed as dist
    torch.set_grad_enabled(False)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = OtherWorkload(args.workload, args, dev)
    B = W.B
    coords = B * W.n_per_item
    planes = W.slice(W.host, 0, B, dev)
    torch.cuda.synchronize()
    fn = lambda: W.decode(planes)
    sampler = ClockSampler(local) if rank == 0 else None   # before the warm-up: nvidia-smi's start-up stalls the device briefly
    for _ in range(args.warmup):
        fn()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        fn()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if sampler else None

    # ---- end to end: pinned host planes -> HBM, decode through the public call, result -> pinned host; items walked in chunks
    CH = max(1, min(args.e2e_chunk, B))
    shape, dt = W.out_item
    out_host = torch.empty((B,) + shape, dtype=dt).pin_memory()
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def e2e_step():
        main = torch.cuda.current_stream()
        chunks = [(k, min(k + CH, B)) for k in range(0, B, CH)]
        ups = []
        with torch.cuda.stream(up):
            for a, b in chunks:
                dp = W.slice(W.host, a, b, dev)
                ev = torch.cuda.Event()
                ev.record(up)
                ups.append((dp, ev))
        for (a, b), (dp, ev) in zip(chunks, ups):
            main.wait_event(ev)
            o = W.decode(dp)
            down.wait_stream(main)
            with torch.cuda.stream(down):
                out_host[a:b].copy_(o.reshape((b - a,) + shape), non_blocking=True)
            o.record_stream(down)
        main.wait_stream(down)
        main.synchronize()

    for _ in range(2):              # two untimed passes: the caching allocator's per-stream pools settle on the second
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(e2e_steps):
        e2e_step()
    a1.record()
    barrier()
    e2e_ms = a0.elapsed_time(a1)
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return

    tf_peak, _, which = peaks()
    per_gpu = coords * args.steps / (ms * 1e-3)
    value = world * per_gpu
    achieved = per_gpu * FLOP_PER_COORD[W.flop_kind] / 1e12          # roofline: per GPU (the kernel's own rate)
    pi = PREC_INFO[args.precision]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": pi['dtype'], "data": "synthetic",
            "config": {"workload": W.desc, "coords_per_step_per_gpu": coords, "precision": args.precision,
                       "parallelism": f"{world} GPU(s), one batch of items per GPU, no exchange on the path",
                       "l2": "planes %.2f GB per step; every step re-reads them (larger than L2 for batch >= 8)" % (W.h2d_bytes() / 1e9)},
            "e2e": {"value": world * coords * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": W.h2d_bytes(),
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size(), "steps": e2e_steps,
                    "pipeline": f"items walked in chunks of {CH}; uploads / read-backs on side streams"},
            "gpu_launches": args.steps * getattr(W, 'launches', 1),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                         "traffic": None, "peak_source": which + " (bf16 sustained)", "kernel": W.kernel,
                         "note": f"achieved = ALGORITHMIC flops ({FLOP_PER_COORD[W.flop_kind]} / coord, as-written layers); " + pi['note'],
                         "executed_tflops": achieved * pi['executed'], "avg_launch_ms": ms / args.steps}}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n, run, what = W.oracle_sample('cpu', args.cpu_coords)
        best = float('inf')
        for _ in range(2):
            t = time.perf_counter()
            run()
            best = min(best, time.perf_counter() - t)
        line["cpu_baseline"] = {"value": n / best, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"oracle port of the reference decoder, fp32, {what}, best of 2, {cores} torch threads"}
        tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        n, run, what = W.oracle_sample(dev, args.cpu_coords * 4)
        dtg = eager_sync_time(run)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
        line["gpu_eager_baseline"] = {"value": n / dtg, "unit": UNIT, "kind": "oracle restatement of the reference decoder, eager "
                                      "PyTorch fp32 (TF32 off) on this GPU", "sample": what + ", best of 3",
                                      "speedup_of_value": per_gpu / (n / dtg)}
    print(json.dumps(line))


def run_mesh(args):
    """Occupancy post-step: one mesh = the reference's dense-grid generation of one decoded latent
    (generation.py:84-98,123-186) -- 128^3 queries in eval_points chunks of 100k, then marching cubes -- everything on the
    device.  metric = meshes / s; the marching-cubes kernels are HBM / latency-bound index work, so the roofline leg is their
    algorithmic bytes over the measured copy bandwidth.  CPU baseline = the reference's own post-step: logits -> host ->
    libmcubes (oracle/_ref), the decode excluded."""
    import numpy as np
    import ddmi_b200
    from ddmi_b200 import generation as gen
    torch.set_grad_enabled(False)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(777)
    torch.manual_seed(777)
    B = args.batch if args.batch != 64 else 8
    m = ddmi_b200.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256).to(dev)
    for blk in (m.net_res1, m.net_res2, m.net_res3, m.net_res4):
        torch.nn.init.kaiming_uniform_(blk.fc_1.weight, a=5 ** 0.5)
    m.precision = args.precision
    host = [[(3.0 * torch.randn(B, 64, s, s, generator=g)).pin_memory() for s in (16, 32, 64)] for _ in range(3)]
    planes = [[t.to(dev) for t in axis] for axis in host]
    item = lambda P, i: tuple([t[i:i + 1] for t in axis] for axis in P)
    nx = 128

    def step(P=planes):
        out = []
        for i in range(B):
            out.append(gen.generate_mesh(item(P, i), m, resolution0=nx))
        return out
    sampler = ClockSampler(0)
    for _ in range(args.warmup):
        res = step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        res = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop()
    # marching cubes alone on the resident grids
    grids = [r[2] for r in res]
    for gr in grids:
        gen.extract_mesh(gr)
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for gr in grids:
        gen.extract_mesh(gr)
    m1.record()
    torch.cuda.synchronize()
    mc_ms = m0.elapsed_time(m1) / B
    nv = sum(r[0].shape[0] for r in res) / B
    nt = sum(r[1].shape[0] for r in res) / B
    # end to end: pinned host planes -> device, mesh -> pinned host
    def e2e():
        P = [[t.to(dev, non_blocking=True) for t in axis] for axis in host]
        return [(v.cpu(), t.cpu()) for v, t, _ in step(P)]
    e2e()
    torch.cuda.synchronize()
    a = time.perf_counter()
    out = e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - a
    _, hbm_peak, which = peaks()
    cells = (nx + 1) ** 3
    alg_bytes = 2 * 4 * nx ** 3 + cells * (4 + 8 + 8 + 4 + 8) + nv * 24 + nt * 24   # grid twice, info + counts + scan + re-reads, outputs
    line = {"metric": "meshes extracted / s (128^3 occupancy grid: eval_points + marching cubes)", "value": B * args.steps / (ms * 1e-3),
            "unit": "meshes/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 vertices / int64 indices (decode: " + args.precision + ")",
            "data": "synthetic",
            "config": {"workload": f"occupancy post-step: {B} latents (triplanes 16^2/32^2/64^2 x64ch), each 128^3 queries in 100k chunks "
                                   f"+ marching cubes at logit(0.2), padded with -1e6", "mean_vertices": nv, "mean_triangles": nt,
                       "l2": "logit grid 8.4 MB + 34 MB of per-cell bookkeeping per mesh: L2-resident"},
            "e2e": {"value": B / e2e_s, "unit": "meshes/s", "h2d_bytes_per_step": sum(t.numel() * 4 for ax in host for t in ax),
                    "d2h_bytes_per_step": int(sum(v.numel() * 8 + t.numel() * 8 for v, t in out)), "steps": 1},
            "gpu_launches": args.steps * B * (21 + 5), "clocks": clocks,
            "breakdown_ms_per_mesh": {"eval_points_plus_marching_cubes": ms / args.steps / B, "marching_cubes": mc_ms},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (mc_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_bytes / (mc_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": which,
                         "kernel": "mcubes::classify / scan / emit (5 launches + one 16-byte read-back per mesh)",
                         "note": "small, launch- and latency-bound: 2.1 M cells per mesh"}}
    if not args.no_cpu_baseline:
        from oracle import mcubes_oracle as mo      # the checker, timed as the baseline only
        gr = grids[0]
        t = time.perf_counter()
        h = gr.cpu().numpy().astype(np.float64)
        refm = mo.extract_mesh(h, 0.2, 0.1, mc=mo.ref_marching_cubes) if mo.ref_marching_cubes(np.zeros((2, 2, 2)), 0.5) is not None else None
        dt = time.perf_counter() - t
        if refm is not None:
            same = bool(np.array_equal(refm[0], res[0][0].cpu().numpy()) and np.array_equal(refm[1], res[0][1].cpu().numpy()))
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "meshes/s (post-step only: D2H + libmcubes)", "cores": 1, "kind": "reference",
                                    "sample": "one 128^3 grid: .cpu() + the reference's marching cubes (oracle/_ref) + vertex normalisation",
                                    "identical_to_gpu_mesh": same, "gpu_post_step_speedup": dt / (mc_ms * 1e-3)}
    print(json.dumps(line))


def run_planes(args):
    """Plane-producer tail at the AFHQ decoder's shapes (configs/d2c-vae/afhq.yaml:27-42: ch 128, ch_mult
    [1,2,4], hdbf at 128 and 64, out_ch 64): heads 512 -> 64 @64^2 and 256 -> 64 @128^2, tail GroupNorm + swish + 3x3 conv
    128 -> 64 @256^2.  metric = plane sets / s.  fp32 CUDA-core kernels: compared with eager PyTorch (cuDNN / cuBLAS fp32, TF32 off)
    on the same GPU and with the oracle on the host cores."""
    import ddmi_b200
    torch.set_grad_enabled(False)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(777)
    B = args.batch if args.batch != 64 else 16
    t = ddmi_b200.PlaneTail(128, 64, (None, 256, 512)).to(dev)
    hs = [torch.randn(B, 512, 64, 64, generator=g).to(dev), torch.randn(B, 256, 128, 128, generator=g).to(dev),
          torch.randn(B, 128, 256, 256, generator=g).to(dev)]

    def step(channels_last=False):
        return [t.head(2, hs[0], channels_last), t.head(1, hs[1], channels_last), t.tail(hs[2], channels_last)]
    sampler = ClockSampler(0)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30                                      # a step is ~6 ms: thirty per "step"
    t0.record()
    for _ in range(args.steps * reps):
        out = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / (args.steps * reps)
    clocks = sampler.stop()
    flops = 2 * B * (64 * 64 * 512 * 64 + 128 * 128 * 256 * 64 + 256 * 256 * 128 * 9 * 64)
    bytes_ = 4 * sum(h.numel() for h in hs) + 4 * sum(o.numel() for o in out)
    tf_peak, hbm_peak, which = peaks()
    line = {"metric": "plane sets emitted / s (AFHQ decoder tail: 2 hdbf heads + norm_out/swish/conv_out)", "value": B / (ms * 1e-3),
            "unit": "plane sets/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (CUDA cores)", "data": "synthetic",
            "config": {"workload": f"plane-producer tail, batch {B}: heads 512->64 @64^2, 256->64 @128^2, tail 128->64 3x3 @256^2",
                       "l2": "inputs %.2f GB per step, larger than L2" % (4 * sum(h.numel() for h in hs) / 1e9)},
            "gpu_launches": 4 * args.steps * reps, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": bytes_ / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_ / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": which,
                         "kernel": "ptail::plane_conv_kernel<3, 64, true> (+ gn_stats_kernel, two 1x1 heads)",
                         "note": "algorithmic bytes = feature maps once + planes once; the 3x3 tail is fp32-FMA-bound, not HBM-bound: "
                                 "%.1f TFLOP/s of fp32 FMA on the CUDA cores (a tcgen05 implicit GEMM is the next step; the tail is "
                                 "~2 %% of a generation step)" % (flops / (ms * 1e-3) / 1e12)}}
    if not args.no_cpu_baseline:
        from oracle import plane_tail_oracle as po      # the checker, timed as the baseline only
        tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        sd = {k: v.detach() for k, v in t.state_dict().items()}
        eager = lambda: [po.head(sd, 2, hs[0]), po.head(sd, 1, hs[1]), po.tail(sd, hs[2], 32, False)]
        ref = eager()
        dtg = eager_sync_time(eager)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
        line["gpu_eager_baseline"] = {"value": B / dtg, "unit": "plane sets/s", "kind": "the oracle (torch conv2d / group_norm, fp32, TF32 off) on this GPU",
                                      "max_abs_vs_ours": max(float((a - b).abs().max()) for a, b in zip(ref, out))}
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sdc = {k: v.cpu() for k, v in sd.items()}
        hc = [h[:1].cpu() for h in hs]
        tt = time.perf_counter()
        po.head(sdc, 2, hc[0]); po.head(sdc, 1, hc[1]); po.tail(sdc, hc[2], 32, False)
        line["cpu_baseline"] = {"value": 1.0 / (time.perf_counter() - tt), "unit": "plane sets/s", "cores": cores, "kind": "port",
                                "sample": "one item, the oracle's torch ops on the host cores"}
    print(json.dumps(line))


def cpu_baseline(res, batch, repeats=1):
    """The oracle port of the reference decoder on the host cores (bounded sample of the same workload)."""
    from oracle import ddmi_oracle as orc   # the checker, timed here as the CPU baseline only
    from ddmi_b200 import convert_to_coord_format_2d, get_scale_injection
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = state_dict32(build_mlp())
    g = torch.Generator().manual_seed(777)
    planes = [torch.randn(batch, 64, s, s, generator=g) for s in (64, 128, 256)]
    e = (res - 1) / res
    coords = convert_to_coord_format_2d(1, res, res, hstart=-e, hend=e, wstart=-e, wend=e)
    si = get_scale_injection(res)
    best = float('inf')
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        orc.image_decode(sd, coords, planes, si)
        best = min(best, time.perf_counter() - t0)
    return {"value": batch * res * res / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port of models/d2c_vae/mlp.py MLP.forward, fp32, batch {batch} @ {res}x{res} "
                      f"({batch * res * res} coords), best of {repeats + 1}, {cores} torch threads"}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', 1))
    n = args.cpu_batch * args.cpu_res * args.cpu_res
    from oracle import ddmi_oracle as orc   # the checker, timed here as the CPU baseline only
    from ddmi_b200 import convert_to_coord_format_2d, get_scale_injection
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = state_dict32(build_mlp())
    g = torch.Generator().manual_seed(777)
    planes = [torch.randn(args.cpu_batch, 64, s, s, generator=g) for s in (64, 128, 256)]
    e = (args.cpu_res - 1) / args.cpu_res
    coords = convert_to_coord_format_2d(1, args.cpu_res, args.cpu_res, hstart=-e, hend=e, wstart=-e, wend=e)
    si = get_scale_injection(args.cpu_res)
    for _ in range(args.warmup):
        orc.image_decode(sd, coords, planes, si)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.image_decode(sd, coords, planes, si)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = (f"oracle port of the reference image decoder (the reference is Python and absent on the GPU box), "
              f"each step = batch {args.cpu_batch} @ {args.cpu_res}x{args.cpu_res} ({n} coords) of the configs[1] workload, "
              f"{cores} torch threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AFHQ-shape image D2C-VAE decode (BASELINE configs[1]), bounded CPU sample", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--res', type=int, nargs='+', default=[1024, 2048])
    ap.add_argument('--store', default='f32', choices=['f32', 'clamp', 'u8'], help='output store mode of the image decoder')
    ap.add_argument('--e2e-chunk', type=int, default=8, help='batch items per pipelined chunk of the e2e leg')
    ap.add_argument('--precision', default='f16f8', choices=['f16f8', 'bf16x3', 'fp32'])
    ap.add_argument('--cpu-res', type=int, default=512)
    ap.add_argument('--cpu-batch', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the strong-scaling (sharded batch 64 + all-gather) leg')
    ap.add_argument('--workload', default='image', choices=['image', 'c1', 'occupancy', 'video', 'nerf', 'mesh', 'planes'],
                    help='image = the headline (BASELINE configs[1]); c1 / video / occupancy / nerf = configs[0], [2], [3], [4] (1 GPU, or N under torchrun: weak scaling)')
    ap.add_argument('--cpu-coords', type=int, default=131072, help='coordinates in the CPU-baseline sample of the non-headline workloads')
    ARGS = ap.parse_args()
    if ARGS.impl == 'reference':
        run_reference(ARGS)
    elif ARGS.workload == 'mesh':
        run_mesh(ARGS)
    elif ARGS.workload == 'planes':
        run_planes(ARGS)
    elif ARGS.workload != 'image':
        run_other(ARGS)
    else:
        run_ours(ARGS)
