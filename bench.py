#!/usr/bin/env python3
"""bench.py -- ORB front-end throughput on B200, one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Metric (BASELINE.json): ORB frames/s on 1241x376 frames, 2000 features, 8 levels.
Headline workload at every N = BASELINE config[1]: 64-frame synthetic stereo batches (32 L/R pairs) per GPU per step;
frames are independent, so ranks share nothing on the data path (weak scaling: each rank extracts its own 64-frame
batch; value = frames of all ranks / max-over-ranks time).
  value     : device-resident -- the step's 64 frames are already in HBM (rotating over a pool of batches larger than the
              126 MB L2), timed with CUDA events on the launching stream.
  e2e       : through the C ABI entry points a caller uses with pinned HOST buffers, H2D of the frames and D2H of
              keypoints + descriptors inside the timed region: orbx_extract_batch_async / orbx_wait with two calls in
              flight (call k+1 submitted, then call k collected); the synchronous orbx_extract_batch is reported beside it.
  strong    : the same 64 frames per step SHARED by the ranks (64 / N per GPU): BASELINE config[1] read literally.
  sustained : the device-resident step looped for >= 2 s (thermal / power steady state), clocks sampled over it.
  configs   : BASELINE configs[2] (1920x1080, 4000 features, 8 levels, 16 frames per step) and configs[3] (3840x2160,
              8000 features, 12 levels, 4 frames per step), each with device-resident and e2e rates and its own roofline.
  matching  : Hamming kNN-2 2000 x 100000 (BASELINE configs[4]), query-sharded over the ranks, gathered everywhere.
  parity_checked : OUTSIDE the timed regions the outputs of the timed workloads are compared with the CPU oracle (sampled
              frames / queries) and the gathered N-rank match records with the single-GPU answer; a mismatch aborts the run.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"),):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# BASELINE.json configs[1..3]; P = sum of level pixels; batch = frames per GPU per step
CONFIGS = {
    "kitti": dict(w=1241, h=376, nf=2000, nl=8, batch=64, P=1444097, pool=6,
                  label="synthetic stereo pairs 1241x376, 2000 features, 8 levels, 64-frame batch (BASELINE config[1])"),
    "hd": dict(w=1920, h=1080, nf=4000, nl=8, batch=16, P=6419321, pool=4,
               label="synthetic 1920x1080 frames, 4000 features, 8 levels, 16-frame batch (BASELINE config[2])"),
    "uhd": dict(w=3840, h=2160, nf=8000, nl=12, batch=4, P=26804551, pool=4,
                label="synthetic 3840x2160 frames, 8000 features, 12 levels, 4-frame batch (BASELINE config[3])"),
}
W, H, NFEAT, NLEVELS, BATCH = 1241, 376, 2000, 8, 64
PIPE_DEPTH = 4          # device-resident batches in flight per GPU (orbx_pipe)
NQ, NT = 2000, 100000
REFERENCE_IMPL = ("the reference's own src/orbextractor.cpp, compiled unmodified, linked to the repo's SCALAR restatement of the "
                  "OpenCV primitives (oracle/cvshim: resize, FAST, GaussianBlur) -- not to OpenCV's SIMD code; a real OpenCV 3.3.1 "
                  "build is roughly 2x faster on the same cores")


def b_alg(cfg):
    """algorithmic bytes per frame, SURVEY 8(d): read the frame, write every pyramid level once, 60 bytes per feature"""
    return cfg["w"] * cfg["h"] + cfg["P"] + 60 * cfg["nf"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML polled every 2 ms from a thread
    (nvidia-smi -lms as the fallback when NVML is not importable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nv, self.stop_flag, self.thread = None, False, None
        self.sm, self.power, self.reasons_seen, self.sm_max = [], [], set(), None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        self.power.append(N.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        r = int(get_reasons(h))
                        for n, bit in names.items():
                            if r & bit:
                                self.reasons_seen.add(n)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nv = N
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
            return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.sm_max,
                    "power_w_max": float(max(self.power)) if self.power else None,
                    "reasons": sorted(self.reasons_seen), "samples": len(self.sm), "source": "NVML polled every 2 ms over the timed regions"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU side (the checker and the reported baseline, never the product path): the reference's own orbextractor.cpp
# (oracle/_ref) when it was built, else the C port
# --------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orb_oracle_py as O
        self.O = O
        ref = os.path.join(ROOT, "oracle", "_ref", "liborbref.so")
        self.kind = "port"
        if os.path.exists(ref):
            try:
                R = C.CDLL(ref)

                class Cfg(C.Structure):
                    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
                R.orbref_create.restype = C.c_void_p
                R.orbref_create.argtypes = [C.POINTER(Cfg)]
                R.orbref_destroy.argtypes = [C.c_void_p]
                R.orbref_run.restype = C.c_int
                R.orbref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
                self.R, self.Cfg, self.kind = R, Cfg, "reference"
            except OSError:
                pass

    def extract_frames(self, frames, threads, nfeat=NFEAT, nlevels=NLEVELS):
        """frames/s over `frames` with `threads` host threads, one extractor instance per thread."""
        O = self.O
        cap = nfeat + 64 * nlevels
        n = len(frames)
        h, w = frames[0].shape
        threads = max(1, min(threads, n))
        if self.kind == "reference":
            def work(t):
                hd = self.R.orbref_create(C.byref(self.Cfg(nfeat, 1.2, nlevels, 20, 7)))
                kps = np.zeros(cap, O.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
                tot = 0
                for i in range(t, n, threads):
                    f = frames[i]
                    tot += self.R.orbref_run(hd, f.ctypes.data, w, h, f.strides[0], kps.ctypes.data, desc.ctypes.data, cap)
                self.R.orbref_destroy(hd)
                return tot
            t0 = time.perf_counter()
            ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            [t.start() for t in ths]; [t.join() for t in ths]
            return n / (time.perf_counter() - t0)
        ex = O.Extractor(nfeat, 1.2, nlevels)
        t0 = time.perf_counter()
        ex.extract_batch_mt(frames, threads)
        return n / (time.perf_counter() - t0)

    def knn2(self, q, t, threads):
        t0 = time.perf_counter()
        self.O.knn2(q, t, nthreads=threads)
        return len(q) * len(t) / (time.perf_counter() - t0)

    @staticmethod
    def cv2_primitives_ms(frame, nlevels):
        """What OpenCV's own (SIMD) resize + FAST + GaussianBlur cost for one frame's pyramid on one host thread: the part of
        the reference's time that the scalar shim of the CPU arm overstates.  None when cv2 is not importable."""
        try:
            import cv2
        except Exception:
            return None
        cv2.setNumThreads(1)
        fast = cv2.FastFeatureDetector_create(7, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

        def once():
            img = frame
            for l in range(nlevels):
                if l:
                    sc = 1.2 ** l
                    img = cv2.resize(img, (int(round(frame.shape[1] / sc)), int(round(frame.shape[0] / sc))), interpolation=cv2.INTER_LINEAR)
                fast.detect(img)
                cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        once()
        t0 = time.perf_counter()
        for _ in range(3):
            once()
        return (time.perf_counter() - t0) / 3 * 1e3


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth_frames(name, rank=0):
    """the frames of one step of a configuration (deterministic; SURVEY 8(d) generators)"""
    import synth
    cfg = CONFIGS[name]
    if name == "kitti":
        return synth.stereo_batch(2 + rank, cfg["w"], cfg["h"], cfg["batch"] // 2)
    base = [synth.scene_s1(cfg["w"], cfg["h"], 1000 * (3 if name == "hd" else 4) + 10 * rank + i) for i in range(min(cfg["batch"], 4))]
    return [np.roll(base[f % len(base)], 5 * (f // len(base)), axis=1) if f >= len(base) else base[f] for f in range(cfg["batch"])]


def headline_pool():
    cfg = CONFIGS["kitti"]
    return max(cfg["pool"], -(-140_000_000 // (BATCH * W * H)))


def headline_config(world, pool, numa):
    """`config` of the headline line; the reference arm prints the same dictionary (it runs the GPU arm's workload)."""
    return {"workload": CONFIGS["kitti"]["label"] + ", one such batch per GPU per step",
            "frames_per_gpu_per_step": BATCH, "global_frames_per_step": BATCH * world,
            "in_flight": f"{PIPE_DEPTH} batches per GPU (orbx_pipe_submit / orbx_pipe_join: consecutive batches on separate extractor "
                         "handles, every batch complete and joined into the timed stream before the closing event)",
            "l2": f"inputs rotate over {pool} resident batches ({pool * BATCH * W * H / 1e6:.0f} MB) + 2x{BATCH * 2.2:.0f} MB of pyramid/blur slabs rewritten every step: working set > 126 MB L2",
            "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
            "host_binding": None if numa is None else f"each rank bound to the {numa} host cores NVML lists for its GPU"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import synth
    cpu = CpuReference()
    cores = host_cores()
    frames = synth_frames("kitti")
    for _ in range(args.warmup):
        cpu.extract_frames(frames[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.extract_frames(frames, cores)
    dt = time.perf_counter() - t0
    fps = BATCH * args.steps / dt
    q, t = synth.matching_set(NQ, NT)
    pairs = cpu.knn2(q[:200], t, cores)
    sample = f"{BATCH}-frame synthetic stereo batch per step on {cores} host threads, one extractor instance per thread"
    other = {}
    for name in ("hd", "uhd"):
        cfg = CONFIGS[name]
        fr = synth_frames(name)
        other[name] = {"value": cpu.extract_frames(fr, cores, cfg["nf"], cfg["nl"]), "unit": "frames/s", "cores": min(cores, len(fr)),
                       "sample": f"one {len(fr)}-frame step of {cfg['label']}, one extractor instance per thread"}
    line = {"impl": "reference", "metric": "ORB frames/s (1241x376, 2000 kp)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": headline_config(args.gpus, headline_pool(), None),      # the GPU arm's own config, key for key
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu.kind, "sample": sample,
                             "reference_impl": REFERENCE_IMPL,
                             "cv2_primitives_ms_per_frame_1thread": cpu.cv2_primitives_ms(frames[0], NLEVELS)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "configs": other,
            "matching": {"value": pairs, "unit": "pairs/s", "cores": cores, "sample": "200 of 2000 queries x 100000 train"}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(index):
    """Run this rank on the host cores next to its GPU (NVML's ideal CPU set) before any pinned buffer is allocated, so
    that first touch puts the staging memory on the GPU's own NUMA node: with 8 ranks the H2D streams otherwise cross the
    socket interconnect.  Returns the number of cores bound to, or None when NVML has no answer."""
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = N.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, v in enumerate(mask) for b in range(64) if (int(v) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ParityError(RuntimeError):
    pass


def check_frames(oex, frames, picks, kps, desc, counts, what):
    """THE CHECKER (outside every timed region): frames[picks] through the CPU oracle against the GPU's records.
    Bit-exact on every field and every descriptor byte; returns the number of key points compared."""
    total = 0
    for f in picks:
        okps, odesc = oex.extract(frames[f])
        n = int(counts[f])
        if n != len(okps):
            raise ParityError(f"{what}: frame {f} has {n} key points, the oracle {len(okps)}")
        g = kps[f][:n]
        for fld in ("x", "y", "size", "angle", "response"):
            if not np.array_equal(g[fld].view(np.uint32), okps[fld].view(np.uint32)):
                raise ParityError(f"{what}: frame {f} field {fld} differs from the oracle")
        for fld in ("octave", "class_id"):
            if not np.array_equal(g[fld], okps[fld]):
                raise ParityError(f"{what}: frame {f} field {fld} differs from the oracle")
        flips = int(np.unpackbits(desc[f][:n] ^ odesc).sum())
        if flips:
            raise ParityError(f"{what}: frame {f} has {flips} descriptor bit flips")
        total += n
    return total


def measure_config(name, args, rank, local_rank, world, torch, dist, orbx, dev, stream, barrier, oracle_mod, frames_per_rank=None,
                   with_sync=True, with_profile=True, sustained_s=0.0, check=8):
    """Device-resident and end-to-end rate of one configuration on this rank's GPU (+ parity of what was timed)."""
    cfg = CONFIGS[name]
    w, h, nf, nl = cfg["w"], cfg["h"], cfg["nf"], cfg["nl"]
    batch = frames_per_rank or cfg["batch"]
    # distinct resident batches: their frames alone exceed the 126 MB L2 (a strong-scaled share is a few MB per step)
    pool = max(cfg["pool"], -(-140_000_000 // (batch * w * h)))
    base = synth_frames(name, rank)
    nb = len(base)
    host_pool, dev_pool, frames_of = [], [], []
    for p in range(pool):
        hb = torch.empty((batch, h, w), dtype=torch.uint8).pin_memory()
        fr = []
        for f in range(batch):
            src = base[(f + 7 * p + batch * rank * (frames_per_rank is not None)) % nb]
            img = np.roll(src, 3 * p, axis=1) if p else src
            hb[f] = torch.from_numpy(img)
            fr.append(img)
        host_pool.append(hb)
        frames_of.append(fr)
        dev_pool.append(hb.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    ex = orbx.Extractor(nf, 1.2, nl, 20, 7, max_width=w, max_height=h, max_batch=batch, device=local_rank)
    cap = ex.max_keypoints

    def step_dev(i):
        ex.extract_batch_device(dev_pool[i % pool].data_ptr(), h * w, w, batch, w, h, stream.cuda_stream)

    for i in range(args.warmup):
        step_dev(i)
    launches_per_step = ex.last_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    res = {"ms_dev": ms_dev, "batch": batch, "launches_per_step": launches_per_step, "cap": cap, "pool": pool}

    # ---- parity of the batch that was just timed (its last step), sampled frames through the oracle
    oex = oracle_mod.Extractor(nfeatures=nf, nlevels=nl)
    parity = {}
    last = (args.steps - 1) % pool
    picks = sorted(set(int(v) for v in np.linspace(0, batch - 1, min(check, batch))))
    kps, desc, cnt = ex.fetch_results(batch, stream.cuda_stream)
    parity["device_batch_frames"] = len(picks)
    parity["device_batch_keypoints"] = check_frames(oex, frames_of[last], picks, kps, desc, cnt, f"{name} device-resident batch")

    # ---- the same steps with PIPE_DEPTH batches in flight (orbx_pipe: consecutive batches go to separate extractor handles, so
    # the head of one batch -- level-0 copy, the dependent resize launches -- runs under the tail of the one before)
    pipe = orbx.Pipe(PIPE_DEPTH, nf, 1.2, nl, 20, 7, max_width=w, max_height=h, max_batch=batch, device=local_rank)

    def run_pipe(n):
        tk = []
        for i in range(n):
            tk.append(pipe.submit(dev_pool[i % pool].data_ptr(), h * w, w, batch, w, h, stream.cuda_stream))
            if i >= PIPE_DEPTH - 1:
                pipe.join(tk[i - (PIPE_DEPTH - 1)], stream.cuda_stream)       # the stream consumes results in submission order
        for j in range(max(0, n - (PIPE_DEPTH - 1)), n):
            pipe.join(tk[j], stream.cuda_stream)
        return tk

    run_pipe(max(args.warmup, PIPE_DEPTH))
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record(stream)
    tk = run_pipe(args.steps)
    p1.record(stream)
    barrier()
    res["ms_pipe"] = p0.elapsed_time(p1)
    picks_p = sorted(set((f + max(1, batch // (3 * len(picks)))) % batch for f in picks))
    kps, desc, cnt = pipe.extractor(tk[-1]).fetch_results(batch, stream.cuda_stream)
    parity["pipe_batch_frames"] = len(picks_p)
    parity["pipe_batch_keypoints"] = check_frames(oex, frames_of[last], picks_p, kps, desc, cnt, f"{name} pipelined device-resident batch")

    # ---- sustained: the pipelined step for >= sustained_s seconds
    if sustained_s > 0:
        per = max(res["ms_pipe"] / args.steps, 1e-3)
        n_sus = int(sustained_s * 1e3 / per) + 1
        clocks = ClockSampler(local_rank).start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record(stream)
        run_pipe(n_sus)
        s1.record(stream)
        barrier()
        res["sustained"] = {"steps": n_sus, "ms": s0.elapsed_time(s1), "clocks": clocks.stop()}
    pipe.close()

    # ---- e2e through the host entry points (pinned host frames in, keypoints + descriptors out)
    # two sets of pinned result arrays with the library's stride (results are DMA'd straight into them): call k+1 fills one
    # while the caller reads call k's
    outs = [(torch.zeros(batch * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(batch, cap),
             torch.zeros((batch, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(batch, np.int32)) for _ in range(2)]
    host_np = [[hb[f].numpy() for f in range(batch)] for hb in host_pool]
    host_ptrs = [orbx.Extractor.frame_pointers(fr) for fr in host_np]     # the `const uint8_t *const *` a C++ caller passes

    def run_async(n):
        prev = None
        for i in range(n):
            t = ex.extract_batch_async(host_ptrs[i % pool], batch, w, h, w, outs[i & 1])
            if prev is not None:
                ex.wait(prev)
            prev = t
        ex.wait(prev)

    run_async(args.warmup)
    barrier()
    t0 = time.perf_counter()
    run_async(args.steps)
    barrier()
    res["s_e2e"] = time.perf_counter() - t0
    res["launches_e2e"] = ex.last_launches()
    o = outs[(args.steps - 1) & 1]
    res["n_kp"] = int(o[2].sum())
    picks2 = sorted(set((f + max(1, batch // (2 * len(picks)))) % batch for f in picks))      # other frames than the device check took
    parity["e2e_frames"] = len(picks2)
    parity["e2e_keypoints"] = check_frames(oex, frames_of[(args.steps - 1) % pool], picks2, o[0], o[1], o[2], f"{name} e2e (async) results")
    if with_sync:
        for i in range(args.warmup):
            ex.extract_batch_ptrs(host_ptrs[i % pool], batch, w, h, w, outs[0])
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            ex.extract_batch_ptrs(host_ptrs[i % pool], batch, w, h, w, outs[0])
        barrier()
        res["s_e2e_sync"] = time.perf_counter() - t0
    # the PCIe floor of the e2e number: pinned H2D copies of a step's frames by themselves
    # (all ranks copy at the same time, as in the e2e loop: with several ranks on one host this is the host's feed rate)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        dev_pool[0].copy_(host_pool[0], non_blocking=True)
        c0.record(stream)
        for p in range(10):
            dev_pool[p % pool].copy_(host_pool[p % pool], non_blocking=True)
        c1.record(stream)
    torch.cuda.synchronize()
    barrier()
    res["h2d_ms"] = c0.elapsed_time(c1) / 10
    if with_profile:
        step_dev(0)
        res["stages"] = ex.profile_stages(reps=5)
    res["parity"] = parity
    res["ex"], res["host_np"], res["host_ptrs"], res["outs"] = ex, host_np, host_ptrs, outs
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only: no HD / UHD configs, no sustained loop, no next rows")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import orbx
    import shard
    import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborbx has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # the checker (CPU oracle): used OUTSIDE the timed regions only
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orb_oracle_py as oracle_mod
    oracle_mod.lib()

    stream = torch.cuda.Stream(device=dev)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- headline: BASELINE config[1], 64 frames per GPU per step
    head = measure_config("kitti", args, rank, local_rank, world, torch, dist, orbx, dev, stream, barrier, oracle_mod,
                          sustained_s=0.0 if args.quick else 2.0)
    clk = clocks.stop() if rank == 0 else None
    ex, host_np, out = head["ex"], head["host_np"], head["outs"][0]
    cap = head["cap"]
    parity = dict(head["parity"])

    # ---- latency of the call OrbFrame makes: one frame (and one stereo pair) through the host entry point
    lat = {}
    for nb, key in ((1, "single_frame_ms"), (2, "stereo_pair_ms")):
        ptrs = orbx.Extractor.frame_pointers(host_np[0][:nb])
        o = (out[0][:nb], out[1][:nb], out[2][:nb])
        for _ in range(10):
            ex.extract_batch_ptrs(ptrs, nb, W, H, W, o)
        t0 = time.perf_counter()
        for _ in range(100):
            ex.extract_batch_ptrs(ptrs, nb, W, H, W, o)
        lat[key] = (time.perf_counter() - t0) / 100 * 1e3

    # ---- strong scaling of config[1]: the 64 frames of a step shared by the ranks (64 / N per GPU)
    strong = None
    if world > 1:
        per = BATCH // world
        sres = measure_config("kitti", args, rank, local_rank, world, torch, dist, orbx, dev, stream, barrier, oracle_mod,
                              frames_per_rank=per, with_sync=False, with_profile=False, check=2)
        sres["ex"].close()
        strong = sres

    # ---- BASELINE configs[2], [3]
    others = {}
    if not args.quick:
        for name in ("hd", "uhd"):
            r = measure_config(name, args, rank, local_rank, world, torch, dist, orbx, dev, stream, barrier, oracle_mod,
                               with_sync=True, check=2 if name == "hd" else 1)
            r["ex"].close()
            for k in ("ex", "host_np", "host_ptrs", "outs"):
                r.pop(k)
            others[name] = r
            torch.cuda.empty_cache()

    # ---- matching: query-sharded kNN-2, the result records gathered on every rank
    q, t = synth.matching_set(NQ, NT)
    m = orbx.Matcher(max_queries=NQ, max_train=NT, device=local_rank)
    nq_loc = shard.query_block(NQ, world)
    q_loc = np.zeros((nq_loc, 32), np.uint8)
    qlo, qhi = shard.query_range(NQ, rank, world)
    part = q[qlo:qhi]
    q_loc[:len(part)] = part
    dq = torch.from_numpy(q_loc).to(dev)
    dt_ = torch.from_numpy(t).to(dev)
    d_loc = torch.empty((nq_loc, 4), dtype=torch.int32, device=dev)
    d_all = torch.empty((world * nq_loc, 4), dtype=torch.int32, device=dev)

    # N > 1: the gather is fused into the matching kernel -- every rank's kernel stores its records into the result window
    # of every rank through NVLink (CUDA IPC mappings), no collective call (orbm_knn2_sharded); the NCCL all_gather of the
    # separate-kernel path is timed beside it
    if world > 1:
        handle = m.window_create(NQ, world, rank)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        for r in range(world):
            if r != rank:
                m.window_attach_ipc(r, handles[r])
        dist.barrier()

    def step_match():
        if world > 1:
            m.knn2_sharded(dq.data_ptr(), qhi - qlo, qlo, dt_.data_ptr(), NT, stream.cuda_stream)
        else:
            m.knn2_device(dq.data_ptr(), nq_loc, dt_.data_ptr(), NT, d_loc.data_ptr(), stream.cuda_stream)

    def step_match_nccl():
        m.knn2_device(dq.data_ptr(), nq_loc, dt_.data_ptr(), NT, d_loc.data_ptr(), stream.cuda_stream)
        with torch.cuda.stream(stream):
            shard.gather_match_records(d_loc, NQ, out=d_all)

    def time_match(fn, n):
        for _ in range(max(args.warmup, 3)):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        barrier()
        return a.elapsed_time(b)

    msteps = max(args.steps, 20)
    ms_match = time_match(step_match, msteps)
    ms_match_nccl = time_match(step_match_nccl, msteps) if world > 1 else 0.0
    if world > 1:
        m.window_status(stream.cuda_stream)
    # parity of the matching result every rank now holds: against the unsharded single-GPU answer, and against the oracle
    if world > 1:
        got = m.window_fetch(NQ, stream.cuda_stream)                 # what the fused kernels left in THIS rank's window
        if not np.array_equal(got[:, :3], d_all[:NQ].cpu().numpy()[:, :3]):
            raise ParityError("peer-stored match records differ from the NCCL-gathered ones")
    else:
        got = d_loc[:NQ].cpu().numpy()
    dq_all = torch.from_numpy(q).to(dev)
    d_one = torch.empty((NQ, 4), dtype=torch.int32, device=dev)
    m.knn2_device(dq_all.data_ptr(), NQ, dt_.data_ptr(), NT, d_one.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    one = d_one.cpu().numpy()
    if not np.array_equal(got[:, :3], one[:, :3]):
        raise ParityError("gathered query-sharded match records differ from the single-GPU answer")
    pick_q = np.linspace(0, NQ - 1, 64).astype(int)
    oi, o1, o2 = oracle_mod.knn2(q[pick_q], t)
    if not (np.array_equal(got[pick_q, 0], oi) and np.array_equal(got[pick_q, 1], o1) and np.array_equal(got[pick_q, 2], o2)):
        raise ParityError("kNN-2 records differ from the oracle")
    parity["knn2_gathered_vs_single_gpu_queries"] = NQ
    parity["knn2_vs_oracle_queries"] = len(pick_q)
    popc_rate = m.measure_popc()

    # ---- the library's own multi-GPU entry points, one process over all GPUs (the other ranks keep their GPUs idle meanwhile)
    lib_multi = None
    if world > 1 and not args.quick:
        store = dist.distributed_c10d._get_default_store()
        barrier()
        if rank == 0:
            try:
                lib_multi = measure_library_multi(args, world, torch, orbx, oracle_mod, q, t, one)
            finally:
                store.set("orbx_lib_multi_done", "1")
        else:
            store.wait(["orbx_lib_multi_done"])
        barrier()

    # ---- "next" rows, rank 0 only, reported beside the headline
    next_rows = None
    if rank == 0 and not args.quick:
        next_rows, nr_ctx = measure_next_rows(torch, orbx, ex, m, head, dev, stream, local_rank)

    # ---- max over ranks
    named = [("ms_dev", head["ms_dev"]), ("s_e2e", head["s_e2e"]), ("s_e2e_sync", head["s_e2e_sync"]), ("ms_pipe", head["ms_pipe"]),
             ("ms_match", ms_match), ("ms_sus", head["sustained"]["ms"] if "sustained" in head else 0.0), ("h2d_ms", head["h2d_ms"])]
    ms_match_nccl = max_over_ranks([ms_match_nccl])[0]
    named += [("ms_strong", strong["ms_dev"] if strong else 0.0), ("s_strong", strong["s_e2e"] if strong else 0.0),
              ("ms_strong_pipe", strong["ms_pipe"] if strong else 0.0)]
    for name in ("hd", "uhd"):
        o = others.get(name)
        named += [(f"{name}_ms_dev", o["ms_dev"] if o else 0.0), (f"{name}_s_e2e", o["s_e2e"] if o else 0.0),
                  (f"{name}_s_e2e_sync", o["s_e2e_sync"] if o else 0.0), (f"{name}_ms_pipe", o["ms_pipe"] if o else 0.0)]
    mx = dict(zip([k for k, _ in named], max_over_ranks([v for _, v in named])))
    ms_dev, s_e2e, s_e2e_sync, ms_pipe, ms_match, ms_sus = (mx[k] for k in ("ms_dev", "s_e2e", "s_e2e_sync", "ms_pipe", "ms_match", "ms_sus"))
    ms_strong, s_strong, ms_strong_pipe = mx["ms_strong"], mx["s_strong"], mx["ms_strong_pipe"]
    kp_checked = sum(v for k, v in parity.items() if k.endswith("_keypoints"))
    for name in others:
        for k, v in others[name]["parity"].items():
            parity[f"{name}_{k}"] = v
    if strong:
        parity["strong_device_batch_frames"] = strong["parity"]["device_batch_frames"]
        parity["strong_e2e_frames"] = strong["parity"]["e2e_frames"]
    parity["mismatches"] = 0          # any mismatch raised ParityError above
    parity["ranks_checked"] = world

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        steps = args.steps
        frames = BATCH * steps * world
        fps_dev = frames / (ms_pipe * 1e-3)          # the headline: PIPE_DEPTH batches in flight (orbx_pipe)
        fps_one = frames / (ms_dev * 1e-3)           # one orbx_extract_batch_device call at a time
        fps_e2e = frames / s_e2e
        stages = head["stages"]
        dom = max(stages, key=stages.get)
        dom_ms = stages[dom]
        balg = b_alg(CONFIGS["kitti"])
        ach = balg * BATCH / (dom_ms * 1e-3) / 1e9
        traffic, winst_step = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            prof = json.load(open(tp))
            traffic = prof.get("bytes_per_step", {}).get(dom)
            winst_step = sum(prof.get("warp_instructions_per_step", {}).values()) or None
        pairs = NQ * NT * msteps / (ms_match * 1e-3)
        int_peak = 148 * sm_max * 1e6 * popc_rate / 8      # pairs/s at the POPC rate measured on this GPU, SURVEY 8(d)
        h2d_ms = mx["h2d_ms"]                      # slowest rank, all ranks copying at once
        line = {
            "metric": "ORB frames/s (1241x376, 2000 kp)", "value": fps_dev, "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms_pipe / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": headline_config(world, head["pool"], numa),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": BATCH * W * H,
                    "d2h_bytes_per_step": BATCH * cap * 60 + BATCH * 4, "keypoints_last_step": head["n_kp"],
                    "api": "orbx_extract_batch_async + orbx_wait, two calls in flight, pinned host buffers",
                    "sync_value": frames / s_e2e_sync, "sync_api": "orbx_extract_batch (one call at a time)",
                    "h2d_alone_ms_per_step": h2d_ms, "h2d_gbs": BATCH * W * H / (h2d_ms * 1e-3) / 1e9,
                    "h2d_aggregate_gbs": world * BATCH * W * H / (h2d_ms * 1e-3) / 1e9,
                    "h2d_note": "pinned host -> device copies of one step's frames by themselves, every rank copying at the same time, slowest rank",
                    "pcie_floor_frames_per_s": BATCH * world / (h2d_ms * 1e-3),
                    "frac_of_pcie_floor": fps_e2e / (BATCH * world / (h2d_ms * 1e-3))},
            "single_call": {"value": fps_one, "unit": "frames/s", "ms_per_step": ms_dev / steps,
                            "what": "the same batches through orbx_extract_batch_device, one call at a time on one handle (round 1's headline)"},
            "gpu_launches": head["launches_per_step"] * steps,
            "gpu_launches_per_step": {"device_resident": head["launches_per_step"], "e2e": head["launches_e2e"]},
            "latency": dict(lat, note="orbx_extract_batch with 1 / 2 frames of 1241x376, pinned host buffers, H2D + kernels + D2H, mean of 100 calls"),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full of this stage, bytes per 64-frame step)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": balg * BATCH, "kernel_ms": dom_ms,
                         "whole_step_frac": balg * BATCH / (ms_pipe / steps * 1e-3) / 1e9 / peak,
                         # the path is instruction-issue bound (50-100 integer operations per byte): warp instructions of one
                         # step (ncu smsp__inst_executed.sum, profiles/ncu_traffic.json) against 4 issue slots x 148 SMs x clock
                         "issue_roofline": None if not winst_step else {
                             "warp_instructions_per_step": winst_step,
                             "peak_warp_instructions_per_s": 4 * 148 * sm_max * 1e6,
                             "frac": winst_step / (ms_pipe / steps * 1e-3) / (4 * 148 * sm_max * 1e6)}},
            "stages_ms": stages,
            "clocks": clk,
            "parity_checked": dict(parity, keypoints_compared_rank0=kp_checked,
                                   how="outside the timed regions: sampled frames of the timed device-resident batch and of the timed e2e "
                                       "results through the CPU oracle (every field and descriptor byte bit-exact); match records gathered "
                                       "over the ranks against the unsharded single-GPU answer and sampled queries against the oracle"),
            "matching": {"metric": "Hamming kNN-2 pairs/s (2000 x 100000)", "value": pairs, "unit": "pairs/s",
                         "queries_per_s": NQ * msteps / (ms_match * 1e-3), "ms_per_batch": ms_match / msteps,
                         "scaling": "strong",
                         "sharding": (f"{nq_loc} queries per GPU, train replicated; ONE kernel per GPU matches its block and stores the 16-byte records into "
                                      "every rank's result window through NVLink peer mappings (CUDA IPC), flag + wait in the kernel, no collective call "
                                      "(orbm_knn2_sharded)") if world > 1 else "single GPU",
                         "nccl_all_gather_path": None if world == 1 else {
                             "ms_per_batch": ms_match_nccl / msteps, "value": NQ * NT * msteps / (ms_match_nccl * 1e-3), "unit": "pairs/s",
                             "what": "two kernels + dist.all_gather_into_tensor of the records (round 1's path), same queries"},
                         "roofline": {"bound": "int", "achieved": pairs / world, "peak": int_peak, "unit": "pairs/s/GPU",
                                      "frac": pairs / world / int_peak, "popc_per_clk_per_sm": popc_rate,
                                      "peak_source": "148 SM x max SM clock x POPC lanes/clk/SM measured on this GPU (orbm_measure_popc) / 8 POPC per pair"}},
        }
        if "sustained" in head:
            sus = head["sustained"]
            line["sustained"] = {"value": BATCH * sus["steps"] * world / (ms_sus * 1e-3), "unit": "frames/s", "seconds": ms_sus * 1e-3,
                                 "steps": sus["steps"], "ms_per_step": ms_sus / sus["steps"], "clocks": sus["clocks"],
                                 "what": "the pipelined device-resident step of the headline looped back to back"}
        if strong:
            per = BATCH // world
            line["strong"] = {"workload": f"BASELINE config[1] read literally: {BATCH} frames per step shared by {world} GPUs ({per} each)",
                              "value": per * world * steps / (ms_strong_pipe * 1e-3), "unit": "frames/s", "ms_per_step": ms_strong_pipe / steps,
                              "single_call": {"value": per * world * steps / (ms_strong * 1e-3), "ms_per_step": ms_strong / steps},
                              "e2e": {"value": per * world * steps / s_strong, "unit": "frames/s"}, "scaling": "strong"}
        else:
            line["strong"] = {"workload": f"BASELINE config[1] read literally: {BATCH} frames per step on 1 GPU = the headline",
                              "value": fps_dev, "unit": "frames/s", "ms_per_step": ms_pipe / steps, "e2e": {"value": fps_e2e, "unit": "frames/s"},
                              "scaling": "strong"}
        cfgs = {}
        for name in ("hd", "uhd"):
            if name not in others:
                continue
            o, cfg = others[name], CONFIGS[name]
            md1, se, ss, md = (mx[f"{name}_{k}"] for k in ("ms_dev", "s_e2e", "s_e2e_sync", "ms_pipe"))
            b = cfg["batch"]
            st = o["stages"]
            dm = max(st, key=st.get)
            ba = b_alg(cfg)
            a = ba * b / (st[dm] * 1e-3) / 1e9
            cfgs[name] = {"workload": cfg["label"] + ", one such batch per GPU per step", "value": b * steps * world / (md * 1e-3),
                          "unit": "frames/s", "ms_per_step": md / steps,
                          "single_call": {"value": b * steps * world / (md1 * 1e-3), "ms_per_step": md1 / steps},
                          "e2e": {"value": b * steps * world / se, "unit": "frames/s", "sync_value": b * steps * world / ss,
                                  "h2d_bytes_per_step": b * cfg["w"] * cfg["h"], "d2h_bytes_per_step": b * o["cap"] * 60 + b * 4,
                                  "pcie_floor_frames_per_s": b * world / (o["h2d_ms"] * 1e-3)},
                          "roofline": {"bound": "hbm", "kernel": dm, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                                       "algorithmic_bytes_per_launch": ba * b, "kernel_ms": st[dm],
                                       "whole_step_frac": ba * b / (md / steps * 1e-3) / 1e9 / peak},
                          "stages_ms": st, "gpu_launches_per_step": o["launches_per_step"]}
        line["configs"] = cfgs
        line["library_multi_gpu"] = lib_multi
        line["next_rows"] = next_rows
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuReference()
            cores = host_cores()
            base = synth_frames("kitti")
            frames_cpu = base * 4                      # 256 frames
            fps_all = cpu.extract_frames(frames_cpu, cores)
            fps_one = cpu.extract_frames(base[:16], 1)
            cv2_ms = cpu.cv2_primitives_ms(base[0], NLEVELS)
            line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": cpu.kind,
                                    "sample": f"{len(frames_cpu)} frames of the same workload over {cores} host threads (one extractor instance per thread); single thread: {fps_one:.1f} frames/s on 16 frames",
                                    "single_thread": fps_one,
                                    "reference_impl": REFERENCE_IMPL,
                                    "cv2_primitives_ms_per_frame_1thread": cv2_ms,
                                    "cv2_note": "cv2 4.13 resize + FAST(7, nms) + GaussianBlur over the same 8-level pyramid on one thread: the OpenCV-side share of "
                                                f"a real reference build; the arm's own single-thread frame takes {1e3 / fps_one:.1f} ms",
                                    "matching_pairs_per_s": cpu.knn2(q[:160], t, cores)}
            if not args.quick:
                for name in cfgs:
                    cfg = CONFIGS[name]
                    fr = synth_frames(name)
                    cfgs[name]["cpu_baseline"] = {"value": cpu.extract_frames(fr, cores, cfg["nf"], cfg["nl"]), "unit": "frames/s",
                                                  "cores": min(cores, len(fr)), "kind": cpu.kind,
                                                  "sample": f"one {len(fr)}-frame step, one extractor instance per thread"}
                cpu_next_rows(line, cpu, orbx, nr_ctx)
        print(json.dumps(line), flush=True)
    ex.close(); m.close()
    if world > 1:
        dist.destroy_process_group()


def measure_library_multi(args, world, torch, orbx, oracle_mod, q, t, one):
    """The multi-GPU entry points of the library itself, driven by ONE process (rank 0) over all `world` GPUs while the other
    ranks idle: orbx_multi_extract_batch_async (frames of a host batch sharded over the GPUs, one host thread per GPU inside
    the library) and orbm_multi_knn2 (query blocks, records exchanged by the kernels through peer stores)."""
    cfg = CONFIGS["kitti"]
    w, h, nf, nl = cfg["w"], cfg["h"], cfg["nf"], cfg["nl"]
    res = {}
    devices = list(range(world))
    oex = oracle_mod.Extractor(nfeatures=nf, nlevels=nl)
    base = synth_frames("kitti")
    for key, batch in (("weak", BATCH * world), ("strong", BATCH)):
        mx = orbx.MultiExtractor(devices, nf, 1.2, nl, 20, 7, max_width=w, max_height=h, max_batch=batch)
        cap = mx.max_keypoints
        pool = 3
        hosts, frames_of = [], []
        for p in range(pool):
            hb = torch.empty((batch, h, w), dtype=torch.uint8).pin_memory()
            fr = []
            for f in range(batch):
                src = base[(f + 5 * p) % len(base)]
                img = np.roll(src, 2 * p + f // len(base), axis=1)
                hb[f] = torch.from_numpy(img)
                fr.append(img)
            hosts.append(hb); frames_of.append(fr)
        ptrs = [orbx.Extractor.frame_pointers([hb[f].numpy() for f in range(batch)]) for hb in hosts]
        outs = [(torch.zeros(batch * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(batch, cap),
                 torch.zeros((batch, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(batch, np.int32)) for _ in range(2)]

        def run(n):
            prev = None
            for i in range(n):
                tk = mx.extract_batch_async(ptrs[i % pool], batch, w, h, w, outs[i & 1])
                if prev is not None:
                    mx.wait(prev)
                prev = tk
            mx.wait(prev)
        run(args.warmup)
        t0 = time.perf_counter()
        run(args.steps)
        dt = time.perf_counter() - t0
        o = outs[(args.steps - 1) & 1]
        picks = sorted(set(int(v) for v in np.linspace(0, batch - 1, 4)))
        check_frames(oex, frames_of[(args.steps - 1) % pool], picks, o[0], o[1], o[2], f"orbx_multi ({key})")
        res[key] = {"value": batch * args.steps / dt, "unit": "frames/s", "frames_per_call": batch, "ms_per_call": dt / args.steps * 1e3,
                    "parity_frames_checked": len(picks)}
        mx.close()
    mm = orbx.MultiMatcher(devices, NQ, NT)
    mm.set_train(t)
    for _ in range(3):
        got = mm.knn2(q)
    t0 = time.perf_counter()
    for _ in range(20):
        got = mm.knn2(q)
    dt = (time.perf_counter() - t0) / 20
    if not (np.array_equal(got[0], one[:, 0]) and np.array_equal(got[1], one[:, 1]) and np.array_equal(got[2], one[:, 2])):
        raise ParityError("orbm_multi_knn2 differs from the single-GPU answer")
    mm.close()
    res["knn2"] = {"ms_per_call": dt * 1e3, "value": NQ * NT / dt, "unit": "pairs/s",
                   "what": "orbm_multi_knn2: 2000 host queries in, records out, train resident on every GPU; all queries equal to the single-GPU answer"}
    res["what"] = f"one process driving {world} GPUs through orbx_multi_* / orbm_multi_* (host threads and peer windows inside liborbx), pinned host buffers, e2e"
    return res


def measure_next_rows(torch, orbx, ex, m, head, dev, stream, local_rank):
    """bag-of-words descent over one batch's descriptors (N2), stereo matching of a batch's pairs (N1), distinctive descriptors
    of a map's points (N4), SearchByProjection with the frame grid (N3)"""
    next_rows = {}
    rng = np.random.default_rng(99)
    out, cap = head["outs"][0], head["cap"]
    voc = orbx.random_vocabulary(10, 5, seed=1)                       # k = 10, L = 5: 111 111 nodes (ORBvoc.txt is k = 10, L = 6)
    V = orbx.Vocabulary(*voc[:5], voc[5], device=local_rank)
    feats = rng.integers(0, 256, (BATCH * NFEAT, 32), dtype=np.uint8)
    dfe = torch.from_numpy(feats).to(dev)
    dwn = torch.empty((BATCH * NFEAT, 2), dtype=torch.int32, device=dev)
    for _ in range(3):
        V.transform_device(dfe.data_ptr(), 32, BATCH * NFEAT, 4, dwn.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record(stream)
    for _ in range(10):
        V.transform_device(dfe.data_ptr(), 32, BATCH * NFEAT, 4, dwn.data_ptr(), stream.cuda_stream)
    n1.record(stream)
    torch.cuda.synchronize()
    ms = n0.elapsed_time(n1) / 10
    next_rows["bow_transform"] = {"features_per_s": BATCH * NFEAT / (ms * 1e-3), "ms_per_64_frames": ms,
                                  "workload": f"{BATCH} frames x {NFEAT} descriptors, synthetic vocabulary k=10 L=5, levels_up=4, device-resident"}
    V.close()
    # stereo matching of the 32 L/R pairs of a batch (N1): pool 0 holds the pairs in L R L R order
    ex.extract_batch_ptrs(head["host_ptrs"][0], BATCH, W, H, W, out)
    sout = (np.zeros((BATCH // 2, cap), np.float32), np.zeros((BATCH // 2, cap), np.float32))
    orbx.stereo_match_batch(ex, BATCH // 2, 0, 1, 2, 386.1448, 0.5372, out=sout)
    t0 = time.perf_counter()
    for _ in range(5):
        _, _, snl, snm = orbx.stereo_match_batch(ex, BATCH // 2, 0, 1, 2, 386.1448, 0.5372, out=sout)
    dt_s = (time.perf_counter() - t0) / 5
    next_rows["stereo_matches"] = {"pairs_per_s": (BATCH // 2) / dt_s, "ms_per_call": dt_s * 1e3, "matches_per_pair": float(snm.mean()),
                                   "workload": f"{BATCH // 2} stereo pairs of one extracted batch in one orbx_stereo_match_batch call, results to host"}
    npts = 20000
    pool = rng.integers(0, 256, (npts * 4, 32), dtype=np.uint8)
    sizes = rng.integers(2, 25, npts)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    inds = rng.integers(0, len(pool), int(offs[-1])).astype(np.int32)
    m.distinctive(pool, offs, inds)
    t0 = time.perf_counter()
    for _ in range(3):
        m.distinctive(pool, offs, inds)
    dt_d = (time.perf_counter() - t0) / 3
    next_rows["distinctive_descriptors"] = {"points_per_s": npts / dt_d, "ms_per_call": dt_d * 1e3,
                                            "workload": f"{npts} map points with 2..24 observations each, host arrays in and out"}
    # SearchByProjection (N3) with the frame grid on the device: frame 0's key points as OBSERVED map points (as in
    # Tracking::SearchLocalPoints), projected 1.25 px beside themselves into frame 0: ~2000 points against ~2000 key points
    k0 = np.ascontiguousarray(out[0][0, :out[2][0]]); d0 = np.ascontiguousarray(out[1][0, :out[2][0]])
    sfl = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    proj_args = (k0, np.full(len(k0), -1, np.float32), None, d0, (0.0, 0.0, float(W), float(H)), d0, k0["x"] + np.float32(1.25), k0["y"].copy(),
                 k0["octave"].copy(), (np.float32(4.0) * sfl[k0["octave"]]).astype(np.float32))
    obs = np.ones(len(k0), np.uint8)
    m.search_by_projection(*proj_args, mp_observed=obs)
    t0 = time.perf_counter()
    for _ in range(20):
        _, _, pnm = m.search_by_projection(*proj_args, mp_observed=obs)
    dt_p = (time.perf_counter() - t0) / 20
    next_rows["search_by_projection"] = {"map_points_per_s": len(k0) / dt_p, "ms_per_call": dt_p * 1e3, "matches": int(pnm),
                                         "workload": f"{len(k0)} observed map points against one frame of {len(k0)} key points, grid + candidate lists + best/second best + the sequential occupancy rule on the device, host arrays in and out"}
    return next_rows, dict(pool=pool, offs=offs, inds=inds, proj_args=proj_args, obs=obs, k0=k0)


def cpu_next_rows(line, cpu, orbx, ctx):
    """the same next-row workloads on one host core through the oracle's restatements (bounded samples)"""
    O = cpu.O
    voc = orbx.random_vocabulary(10, 5, seed=1)
    fs = np.random.default_rng(99).integers(0, 256, (20000, 32), dtype=np.uint8)
    t0 = time.perf_counter(); O.voc_transform(voc[0], voc[1], voc[2], voc[3], voc[5], 4, fs); dtv = time.perf_counter() - t0
    line["next_rows"]["bow_transform"]["cpu_port_features_per_s_1core"] = len(fs) / dtv
    pool, offs, inds = ctx["pool"], ctx["offs"], ctx["inds"]
    t0 = time.perf_counter(); O.distinctive(pool, offs[:2001], inds[:offs[2000]]); dtd = time.perf_counter() - t0
    line["next_rows"]["distinctive_descriptors"]["cpu_port_points_per_s_1core"] = 2000 / dtd
    t0 = time.perf_counter()
    for _ in range(5):
        O.search_by_projection(*ctx["proj_args"], mp_observed=ctx["obs"])
    line["next_rows"]["search_by_projection"]["cpu_port_map_points_per_s_1core"] = len(ctx["k0"]) * 5 / (time.perf_counter() - t0)


if __name__ == "__main__":
    main()
