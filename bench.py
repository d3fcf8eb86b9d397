#!/usr/bin/env python3
"""bench.py — ORB extract+describe frames/s on the BASELINE.json headline config (752x480, 1000 key points, euroc settings
1.2 / 8 levels / FAST 20,7), plus all-pairs Hamming GMatch/s, on N B200s of one node.

A step = one pass of the hot path (pyramid, FAST, quadtree, blur, descriptors) over one batch of B synthetic frames per GPU.
  value : frames/s with the frames already resident in HBM (device API), CUDA events on the launching stream, max over ranks
  e2e   : frames/s through the host C-ABI with pinned HOST buffers — H2D and D2H of every step inside the timed region; the streaming
          form orbfe_extract_batch_submit / _wait with two batches in flight (e2e.value) and one blocking orbfe_extract_batch call
          per step (e2e.sync_call_value)
  roofline : the dominant stage, algorithmic bytes per launch / its CUDA-event duration, against MEASURED_PEAKS.json
  cpu_baseline : the reference's own ORBExtractor.cpp (oracle/_ref, compiled verbatim) on the host cores, bounded sample
`--impl reference` runs only that CPU arm.  N > 1: frames are sharded by rank (weak scaling), results gathered with NCCL."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NF = 752, 480, 1000
ORB = dict(nFeatures=NF, scaleFactor=1.2, nLevels=8, iniThFast=20, minThFast=7)
WORKLOAD = "C1: %dx%d grayscale frames, %d features, 8 levels, scale 1.2, FAST 20/7 (euroc-shaped), dense synthetic profile" % (W, H, NF)


def level_sizes(w, h, n_levels=8, sf=1.2):
    sc = [np.float32(1.0)]
    for _ in range(1, n_levels):
        sc.append(np.float32(sc[-1] * np.float32(sf)))
    out = [(w, h)]
    for l in range(1, n_levels):
        inv = np.float32(1.0) / sc[l]
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(w, h, n_kp):
    """per-frame algorithmic bytes of each stage."""
    a = [x * y for x, y in level_sizes(w, h)]
    s = sum(a)
    return {"pyramid": (s - a[-1]) + (s - a[0]), "fast": s, "blur": 2 * s, "describe": 60 * n_kp, "quadtree": 0,
            "frame_total": (s - a[-1]) + (s - a[0]) + s + 2 * s + 60 * n_kp}


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag, self.max_mhz = gpu, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_fps(frames, threads, min_seconds=8.0, max_frames=None):
    """Time the reference's own extractor (oracle/_ref/libref_orb.so = ORBExtractor.cpp compiled verbatim + cv shim) frame-parallel
    on `threads` host threads; falls back to the C oracle port when the verbatim build is not there.  Returns (fps, kind, sample)."""
    from oracle import orb_oracle as orc
    kind = "reference"
    try:
        ref = orc.ReferenceExtractor(NF, 1.2, 8, 20, 7, canonical=False)
    except (FileNotFoundError, OSError):
        ref, kind = None, "port"
    n = max(threads, 8)
    total_t, total_n = 0.0, 0
    while True:
        idx = [i % len(frames) for i in range(total_n, total_n + n)]
        batch = np.ascontiguousarray(frames[idx])
        if ref is not None:
            sec, counts = ref.time_batch(batch, threads)
            assert counts.min() > 0
        else:
            ex = orc.Extractor(NF, 1.2, 8, 20, 7)
            t0 = time.perf_counter()
            for f in batch:
                ex(f)
            sec = time.perf_counter() - t0
        total_t += sec; total_n += n
        if total_t >= min_seconds or (max_frames and total_n >= max_frames):
            break
        n = min(4 * n, max(n, int(n * (min_seconds - total_t) / max(sec, 1e-6)) + threads))
    used = threads if ref is not None else 1
    return total_n / total_t, kind, "%d frames of the workload in %.1f s on %d thread(s)" % (total_n, total_t, used), used


def cpu_hamming_gmatch(threads, nt=40000, q_per_thread=1024):
    """The reference's DescriptorDistance arithmetic (ORBMatcher.cpp:17-31, SWAR popcount; oracle port) in a best/second-best loop,
    one slice of queries per host thread against the 40k-descriptor table."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orb_oracle as orc
    rng = np.random.default_rng(5)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    qs = [rng.integers(0, 256, (q_per_thread, 32), dtype=np.uint8) for _ in range(threads)]
    orc.hamming_allpairs(qs[0][:8], t)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(lambda q: orc.hamming_allpairs(q, t), qs))
    sec = time.perf_counter() - t0
    return {"value": threads * q_per_thread * nt / sec / 1e9, "unit": "GMatch/s", "cores": threads, "kind": "port",
            "sample": "%d x %d pairs in %.2f s on %d thread(s)" % (threads * q_per_thread, nt, sec, threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from monoorbslam3_b200 import synth
    cores = host_cores()
    frames = synth.frames(16, H, W, 1000, "dense")
    for _ in range(max(args.warmup, 0)):
        cpu_reference_fps(frames, cores, min_seconds=0.5)
    vals, samples = [], []
    t_all = time.perf_counter()
    kind = used = None
    for _ in range(args.steps):
        fps, kind, sample, used = cpu_reference_fps(frames, cores, min_seconds=args.ref_seconds)
        vals.append(fps); samples.append(sample)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "ORB extract+describe frames/s (752x480, 1000 kp)", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": "bounded CPU sample of the same frames (%s)" % samples[-1],
                       "sharding": "rank 0 only, all host threads", "l2": "n/a (CPU arm)"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": used, "kind": kind, "sample": samples[-1]},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from monoorbslam3_b200 import ORBExtractor, ORBMatcher, KP_DTYPE, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries exactly one JSON line, so the
        # banner is sent to stderr (fd-level, NCCL writes from C)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    # synthetic frames: `distinct` different scenes, repeated to fill the batch (B*H*W = 185 MB at B=512 > the 126 MB L2)
    distinct = min(args.distinct, B)
    base = synth.frames(distinct, H, W, 1000 + 100000 * rank, "dense")
    reps = (B + distinct - 1) // distinct
    host_frames = torch.from_numpy(np.concatenate([base] * reps)[:B]).pin_memory()
    d_frames = host_frames.to(dev, non_blocking=True)

    ex = ORBExtractor(device=local, max_batch=B, **ORB)
    cap = NF + 64
    # a real (non-default) stream: the library launches on it and the timing events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # Output slabs.  N > 1: two sets, so that the NCCL gather of step k (fixed-capacity slabs) runs on a side stream
    # while step k+1 computes into the other set — the gather is inside the timed region but off the compute stream's critical path.
    n_sets = 2 if world > 1 else 1
    outs = [(torch.zeros((B, cap, 7), dtype=torch.float32, device=dev), torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev),
             torch.zeros(B, dtype=torch.int32, device=dev)) for _ in range(n_sets)]
    d_kps, d_desc, d_n = outs[0]
    gather_bufs = comm_stream = None
    if world > 1:
        gather_bufs = [(torch.empty((world,) + tuple(d_kps.shape), dtype=d_kps.dtype, device=dev),
                        torch.empty((world,) + tuple(d_desc.shape), dtype=d_desc.dtype, device=dev),
                        torch.empty((world, B), dtype=torch.int32, device=dev)) for _ in range(n_sets)]
        comm_stream = torch.cuda.Stream(device=dev)
        ev_computed = [torch.cuda.Event() for _ in range(n_sets)]
        ev_gathered = [torch.cuda.Event() for _ in range(n_sets)]
    step_no = [0]

    def step_device():
        k = step_no[0] % n_sets
        step_no[0] += 1
        o = outs[k]
        if world > 1 and step_no[0] > n_sets:
            tstream.wait_event(ev_gathered[k])             # the gather that read this set two steps ago has finished
        ex.extract_batch_device(d_frames, B, H, W, o[0], o[1], cap, o[2], stream=stream, sync=False)
        if world > 1:
            ev_computed[k].record(tstream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev_computed[k])
                dist.all_gather_into_tensor(gather_bufs[k][0], o[0])
                dist.all_gather_into_tensor(gather_bufs[k][1], o[1])
                dist.all_gather_into_tensor(gather_bufs[k][2], o[2])
                ev_gathered[k].record(comm_stream)

    def drain():
        """the timed region ends when the last gather has landed: the compute stream waits for the side stream"""
        if world > 1:
            tstream.wait_stream(comm_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step_device()
    barrier()
    n_kp_mean = float(d_n.float().mean().item())
    assert n_kp_mean > 900, n_kp_mean

    # ---- timed region 1: frames resident in HBM
    sampler = ClockSampler(local); sampler.start()
    launches0 = ex.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        step_device()
    drain()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = ex.launch_count() - launches0
    # the same K steps again with per-stage CUDA events recorded inside the library on the launching stream; with the events on,
    # the stages run back to back on one stream (the unprofiled run overlaps the blur with FAST + quadtree on a second stream)
    ex.profile(True)
    step_device(); drain()            # untimed: the profiled pass runs the whole batch in one arena (the unprofiled one splits it over two)
    ex.profile_read(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(K):
        step_device()
    drain()
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    stage_ms, passes = ex.profile_read(reset=True)
    ex.profile(False)
    if world > 1:
        t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    value = world * B * K / (ms_dev * 1e-3)

    # ---- timed region 2: end to end through the host C-ABI with pinned host buffers (H2D + kernels + D2H every step)
    h_n = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_kps = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
    frames_np = host_frames.numpy()
    out = (h_n.numpy(), h_kps.numpy().view(KP_DTYPE).reshape(B, cap), h_desc.numpy())
    def timed(run):
        barrier()
        t0 = time.perf_counter()
        run()
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return world * B * K / sec

    # (a) one synchronous orbfe_extract_batch call per step: every call pays the pipeline's fill (first upload) and drain (last pass + download)
    def run_sync():
        for _ in range(K):
            ex.extract_batch(frames_np, cap=cap, out=out)
            _ = int(out[0][0])            # the step's result is read on the host
    for _ in range(2):
        ex.extract_batch(frames_np, cap=cap, out=out)
    e2e_sync = timed(run_sync)

    # (b) the streaming form of the same entry point: step k is submitted before step k-1 is waited for (two batches in flight, two
    # sets of pinned output buffers), so the uploads of step k run under the last passes of step k-1.  Every step still uploads its
    # frames from pinned host memory and has its result read on the host inside the timed region.
    h_n2 = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_kps2 = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
    h_desc2 = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
    outs = (out, (h_n2.numpy(), h_kps2.numpy().view(KP_DTYPE).reshape(B, cap), h_desc2.numpy()))

    def run_stream():
        prev = None
        for k in range(K):
            t = ex.extract_batch_submit(frames_np, outs[k & 1], cap=cap)
            if prev is not None:
                ex.extract_batch_wait(prev)
                _ = int(outs[(k - 1) & 1][0][0])
            prev = t
        ex.extract_batch_wait(prev)
        _ = int(outs[(K - 1) & 1][0][0])
    run_stream()
    e2e = timed(run_stream)
    # both buffer sets hold the results of the same frames (rows beyond n[b] are unspecified)
    assert np.array_equal(outs[0][0], outs[1][0]), "streamed batches disagree on the key-point counts"
    for b in range(B):
        nb_ = int(outs[0][0][b])
        assert np.array_equal(outs[0][2][b, :nb_], outs[1][2][b, :nb_]) and outs[0][1][b, :nb_].tobytes() == outs[1][1][b, :nb_].tobytes()
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- all-pairs Hamming (BASELINE config 4: 20 key frames x 2000 descriptors = 40k x 40k), device resident
    nq = 40000
    g = torch.Generator(device="cpu"); g.manual_seed(7 + rank)
    descs = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g).to(dev)
    bi = torch.zeros(nq, dtype=torch.int32, device=dev); bd = torch.zeros_like(bi); sd = torch.zeros_like(bi)
    mt = ORBMatcher(0.6, False, handle=ex._h)

    def time_allpairs():
        for _ in range(2):
            mt.hamming_allpairs_device(descs, nq, descs, nq, bi, bd, sd, stream=stream, sync=False)
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps_m = 5
        m0.record()
        for _ in range(reps_m):
            mt.hamming_allpairs_device(descs, nq, descs, nq, bi, bd, sd, stream=stream, sync=False)
        m1.record()
        torch.cuda.synchronize()
        return m0.elapsed_time(m1) / reps_m, (bi.clone(), bd.clone(), sd.clone())

    # the int8 tensor-core formulation (default for problems of this size), then the popc kernel on the same inputs
    ms_match, res_tc = time_allpairs()
    gmatch = nq * nq / (ms_match * 1e-3) / 1e9
    os.environ["ORBFE_ALLPAIRS_POPC"] = "1"
    ms_popc, res_popc = time_allpairs()
    del os.environ["ORBFE_ALLPAIRS_POPC"]
    gmatch_popc = nq * nq / (ms_popc * 1e-3) / 1e9
    assert all(torch.equal(a, b) for a, b in zip(res_tc, res_popc)), "tensor-core and popc all-pairs disagree"
    imma_peak = mt.imma_peak()            # measured 10^9 pairs/s of the mma.sync int8 instruction
    popc_peak = mt.popc_peak()            # measured 10^9 popc/s; one match = 8 popc

    # ---- single-frame latency through the reference-shaped call (ORBExtractor::operator(), host image in, host vectors out)
    one = base[0]
    ex1 = ORBExtractor(device=local, max_batch=1, **ORB)
    for _ in range(5):
        ex1(one)
    t0 = time.perf_counter()
    for _ in range(100):
        ex1(one)
    ms_single = (time.perf_counter() - t0) * 1e3 / 100

    # ---- SearchForInitialization on a frame pair of the workload (window 100, ratio 0.9), GPU call vs the CPU restatement
    init = None
    if rank == 0:
        from monoorbslam3_b200 import FrameView
        fa_, fb_ = synth.shifted_pair(H, W, 1000)
        ex2 = ORBExtractor(device=local, max_batch=1, **dict(ORB, nFeatures=2 * NF))       # the initial extractor uses 2 x nFeatures (Tracking.cpp:24)
        ka, da = ex2(fa_); kb, db = ex2(fb_)
        f1, f2 = FrameView(ka, da, W, H), FrameView(kb, db, W, H)
        mi = ORBMatcher(0.9, True, handle=ex2._h)
        pre0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
        for _ in range(3):
            mi.SearchForInitialization(f1, f2, pre0.copy(), 100)
        t0 = time.perf_counter()
        for _ in range(20):
            n_init, _ = mi.SearchForInitialization(f1, f2, pre0.copy(), 100)
        ms_init = (time.perf_counter() - t0) * 1e3 / 20
        init = {"matches": int(n_init), "gpu_ms_per_call": ms_init, "queries": int((ka["octave"] == 0).sum())}

    # ---- the other single-GPU BASELINE configs, frames resident in HBM (reported next to the headline, not the headline)
    other = None
    if rank == 0:
        other = {}
        for name, (ow, oh, onf, ob) in {"C2: 1241x376, 2000 features (kitti-shaped)": (1241, 376, 2000, 256),
                                         "C3: 1920x1080, 4000 features (phone-shaped)": (1920, 1080, 4000, 128)}.items():
            ob_frames = torch.from_numpy(synth.frames(8, oh, ow, 1000, "dense")).to(dev)[torch.arange(ob, device=dev) % 8].contiguous()
            oex = ORBExtractor(device=local, max_batch=ob, **dict(ORB, nFeatures=onf))
            ocap = onf + 128
            okps = torch.zeros((ob, ocap, 7), dtype=torch.float32, device=dev); odesc = torch.zeros((ob, ocap, 32), dtype=torch.uint8, device=dev)
            on_ = torch.zeros(ob, dtype=torch.int32, device=dev)
            for _ in range(3):
                oex.extract_batch_device(ob_frames, ob, oh, ow, okps, odesc, ocap, on_, stream=stream, sync=False)
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(5):
                oex.extract_batch_device(ob_frames, ob, oh, ow, okps, odesc, ocap, on_, stream=stream, sync=False)
            o1.record(); torch.cuda.synchronize()
            oms = o0.elapsed_time(o1) / 5
            other[name] = {"frames_per_s": ob / oms * 1e3, "ms_per_pass": oms, "frames_per_pass": ob, "mean_keypoints_per_frame": float(on_.float().mean().item())}
            oex.close(); del ob_frames, okps, odesc, on_

    # ---- tracking-loop matchers on the same frame pair: SearchByProjection (frame -> frame, th 15) and the local-map search (th 2)
    track = None
    if rank == 0:
        rng = np.random.default_rng(0)
        nq = len(ka)
        q_u = (ka["x"] - 7 + rng.normal(0, 1.0, nq)).astype(np.float32); q_v = (ka["y"] - 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
        q_l = ka["octave"].astype(np.int32); q_a = ka["angle"].astype(np.float32); q_ok = (rng.random(nq) < 0.9).astype(np.uint8)
        occ = np.zeros(len(kb), np.uint8)
        sfac = np.array([ex2.getScaleFactor(int(l)) for l in q_l], np.float32)
        r_proj = (np.float32(15) * ka["size"]).astype(np.float32); r_loc = (np.float32(2) * np.float32(4.0) * sfac).astype(np.float32)
        mt1 = ORBMatcher(0.9, True, handle=ex2._h); mt2 = ORBMatcher(0.8, True, handle=ex2._h)

        def _ms(f, reps):
            for _ in range(3):
                f()
            t0 = time.perf_counter()
            for _ in range(reps):
                r = f()
            return (time.perf_counter() - t0) * 1e3 / reps, r
        g1, (n1_, _) = _ms(lambda: mt1.SearchByProjection(q_u, q_v, r_proj, q_l, q_a, da, q_ok, f2, occ), 20)
        g2, (n2_, _) = _ms(lambda: mt2.SearchLocalPoints(q_u, q_v, r_loc, q_l, da, q_ok, f2, occ), 20)
        track = {"queries": int(nq), "search_by_projection_th15": {"gpu_ms_per_call": g1, "matches": int(n1_)},
                 "search_local_points_th2": {"gpu_ms_per_call": g2, "matches": int(n2_)}}
        if world == 1 and not args.no_cpu:
            from oracle import orb_oracle as orc
            c1, (cn1, _) = _ms(lambda: orc.search_by_projection(q_u, q_v, r_proj, q_l, q_a, da, q_ok, kb, db, W, H, occ, True), 5)
            c2, (cn2, _) = _ms(lambda: orc.search_local_points(q_u, q_v, r_loc, q_l, da, q_ok, kb, db, W, H, occ, 0.8), 5)
            track["search_by_projection_th15"].update(cpu_port_ms_per_call=c1, cpu_matches=int(cn1))
            track["search_local_points_th2"].update(cpu_port_ms_per_call=c2, cpu_matches=int(cn2))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        ab = algorithmic_bytes(W, H, n_kp_mean)
        per_launch_ms = {k: v / max(passes, 1) for k, v in stage_ms.items()}
        hbm_stages = ("pyramid", "fast", "blur")
        dom = max(hbm_stages, key=lambda k: per_launch_ms[k])
        achieved = ab[dom] * B / (per_launch_ms[dom] * 1e-3) / 1e9
        stage_report = {k: {"ms_per_step": per_launch_ms[k], "algorithmic_GBps": (ab[k] * B / (per_launch_ms[k] * 1e-3) / 1e9) if per_launch_ms[k] > 0 else None,
                            "frac_of_hbm_peak": (ab[k] * B / (per_launch_ms[k] * 1e-3) / 1e9 / peak_gbs) if per_launch_ms[k] > 0 and ab[k] else None}
                        for k in stage_ms}
        traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
        issue = None            # why the HBM fraction is low: the stage kernels are bound by instruction issue, not by bytes
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("batch") == B:
                traffic = tj["stages"][dom]["dram_read_bytes"] + tj["stages"][dom]["dram_write_bytes"]
                clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
                slots_per_s = 148 * 4 * clk                       # warp instructions the SM sub-partitions can issue per second
                issue = {}
                for k, d in tj["stages"].items():
                    if "warp_instructions" not in d or per_launch_ms.get(k, 0) <= 0:
                        continue
                    floor_ms = d["warp_instructions"] / slots_per_s * 1e3
                    issue[k] = {"warp_instructions_per_step": d["warp_instructions"], "issue_floor_ms": floor_ms,
                                "issue_frac": floor_ms / per_launch_ms[k], "ncu_issue_active_pct": d.get("issue_active_pct"),
                                "ncu_alu_pipe_pct": d.get("alu_pipe_pct"),
                                "thread_instructions_per_pixel": (d["warp_instructions"] * 32 / (ab["fast"] * B)) if k in ("fast", "blur") else None}
        except Exception:
            pass
        line = {
            "metric": "ORB extract+describe frames/s (752x480, 1000 kp)", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "distinct_scenes": distinct, "sharding": "frames by rank, NCCL all_gather of result slabs on a side stream (double-buffered, inside the timed region)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (%.0f MB of frames per step)" % (B * H * W / 1e6), "mean_keypoints_per_frame": n_kp_mean},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * H * W, "d2h_bytes_per_step": B * (4 + cap * 60),
                    "api": "orbfe_extract_batch_submit / _wait (host C-ABI, pinned host buffers, two batches in flight)",
                    "sync_call_value": e2e_sync, "sync_call_api": "orbfe_extract_batch, one blocking call per step"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab[dom] * B,
                         "whole_step_algorithmic_GBps": ab["frame_total"] * B / (ms_dev / K * 1e-3) / 1e9},
            "stages": stage_report, "ms_per_step_serialised_with_stage_events": ms_prof / K,
            "issue_roofline": {"note": "warp instructions per step (committed ncu capture) / (148 SMs x 4 schedulers x SM clock): the time the stage would "
                                       "take at one instruction per scheduler per cycle; issue_frac = that / measured stage time", "stages": issue},
            "match": {"metric": "Hamming GMatch/s (all-pairs 40000 x 40000, best/second-best)", "value": gmatch, "unit": "GMatch/s", "ms": ms_match,
                      "kernel": "k_allpairs_imma: +-1 int8 GEMM on the tensor cores (mma.sync m16n8k32), hamming = (256 - dot) / 2, fused (min, second-min) epilogue; identical results to the popc kernel (checked in this run)",
                      "roofline": {"bound": "tensor (int8 IMMA via mma.sync)", "achieved": gmatch, "peak": imma_peak, "unit": "GMatch/s", "frac": gmatch / imma_peak,
                                   "peak_source": "measured IMMA micro-benchmark (orbfe_imma_peak), 256 int8 multiply-adds per match"},
                      "popc_kernel": {"value": gmatch_popc, "unit": "GMatch/s", "ms": ms_popc,
                                      "roofline": {"bound": "integer pipe (popc)", "achieved": gmatch_popc, "peak": popc_peak / 8, "unit": "GMatch/s",
                                                   "frac": gmatch_popc / (popc_peak / 8),
                                                   "peak_source": "measured popc micro-benchmark (orbfe_popc_peak), 8 popc per 256-bit match"}}},
            "single_frame": {"ms_per_call": ms_single, "frames_per_s": 1e3 / ms_single, "api": "ORBExtractor.__call__ -> orbfe_extract (host image in, host key points out)"},
            "search_for_initialization": init,
            "tracking_matchers": track,
            "other_configs": other,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            fps, kind, sample, used = cpu_reference_fps(base, host_cores(), min_seconds=args.ref_seconds)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": used, "kind": kind, "sample": sample}
            line["match"]["cpu_baseline"] = cpu_hamming_gmatch(host_cores())
            if init is not None:
                from oracle import orb_oracle as orc
                t0 = time.perf_counter()
                for _ in range(5):
                    on, _, _ = orc.search_for_initialization(ka, da, kb, db, W, H, pre0.copy(), 100, 0.9, True)
                init["cpu_port_ms_per_call"] = (time.perf_counter() - t0) * 1e3 / 5
                init["cpu_matches"] = int(on)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frames per step per GPU")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic scenes per rank (repeated to fill the batch)")
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="CPU work per reference step / cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
