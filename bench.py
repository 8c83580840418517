#!/usr/bin/env python
"""Benchmark of the DH3D hot path (BASELINE.json metric: point-clouds/sec, full DH3D forward,
N = 8192, at 1/2/4/8 B200; HBM GB/s vs roofline).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      CPU reference arm (oracle port, host cores)
    python bench.py --workload local|retrieval|sweep ...      the other BASELINE.json configs

Workloads (BASELINE.json `configs`):
  full       configs[2] (default): full local + detector + global forward, 32 clouds of 8192 points per GPU per
             step, weak scaling; every rank keeps its [32,256] global descriptors per step and ONE NCCL all-gather
             of all of them closes the timed region (north_star: "one all-gather of the global descriptors at the end").
  local      configs[1]: local descriptor forward (basic_config), batch 8.
  retrieval  configs[3]: 4096 clouds sharded over the ranks (512 per GPU at 8 GPUs) in micro-batches of 32, only the
             global branch's output is fetched (what evaluate/global_eval/globaldesc_extract.py does), one all-gather
             of the 256-D descriptors at the end, then the k = 25 retrieval (evaluation_retrieval.py:37-53). Strong scaling.
  sweep      configs[4]: FlexConv + k-NN over N in {4096,8192,16384,32768} x K in {8,16,32}, C = 128, B = 8, 1 GPU:
             achieved algorithmic HBM GB/s against the measured roofline, native and drop-in (reference-layout) entries.

One JSON line on stdout (rank 0).  `value` = clouds of all ranks / max-over-ranks device time with inputs resident
in HBM, two steps in flight per GPU (dh3d_b200.model.InFlightForward: the two CUDA-graph instances of the forward
replay on their own streams; `--in-flight 1` = strictly one step after the other, also reported as
`one_step_in_flight`); `e2e` = the same forward through the public API from pinned HOST buffers with the H2D copy of the clouds and
ONE D2H copy of the step's whole result block inside the timed region, next to the measured pinned-D2H ceiling of
the box; `e2e_modes` = the reference's two other output modes (global descriptors only; --perform_nms keypoint
rows only); `roofline` = the dominant kernel timed live with CUDA events; `op_roofline` / `step_hbm` = every op's
algorithmic bytes against the HBM roof; `sustained` = a >= 2 s loop with its own clock / power record;
`ref_cuda` = the reference's own CUDA kernels (oracle/_ref, comparator only) on the same inputs in the same run;
`data_sensitivity` = uniform / LiDAR-like / real demo / all-zero / far-outlier inputs; `cpu_baseline` = the oracle
port of the reference's CPU path on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point_clouds_per_sec_full_dh3d_forward_n8192"
N_POINTS = 8192


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """SM clock / power / throttle-reason samples DURING a timed region (B200_PROFILING.md): an in-process NVML
    polling thread (2 ms period, so even a 50 ms region gets samples); falls back to `nvidia-smi -lms`."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.samples, self.power, self.mask, self.max_mhz, self.stop_flag, self.nvml = [], [], 0, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    phys = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mask |= int(get_reasons(self.handle))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.samples)
            reasons = sorted(name for name, bit in self.REASONS if self.mask & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                    "sm_max_mhz": self.max_mhz, "samples": len(sm), "reasons": reasons,
                    "power_w_max": max(self.power) if self.power else None,
                    "power_w_mean": (sum(self.power) / len(self.power)) if self.power else None, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], None, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "source": "nvidia-smi"}


def synth_clouds(batch, n_points, seed):
    """xyz ~ U(-25, 25)^3 fp32, torch.Generator().manual_seed(1234 + id)."""
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    pts = (torch.rand((batch, n_points, 3), generator=g) * 50.0 - 25.0).float()
    return pts


# ---- algorithmic bytes / flops per launch group (SURVEY 8(d) formulas), and what bounds the op -------------------
def algorithmic(tag, name):
    """-> (bytes, flops, bound) of one call from its shape tag.  bound: 'hbm' (gather / streaming ops: graded
    against the HBM roof), 'tensor' (dense 1x1 GEMMs), 'alu' (all-pairs distance + selection: k-NN, 3-NN),
    'latency' (FPS: M serial rounds)."""
    d = {}
    for part in tag.split("_"):
        key = part.rstrip("0123456789")
        if key and part[len(key):]:
            d[key] = int(part[len(key):])
    if name in ("dh3d_linear", "dh3d_linear_packed"):
        M, K, N = d["M"], d["K"], d["N"]
        return 4.0 * (M * K + K * N + M * N), 2.0 * M * K * N, "tensor"
    if name == "dh3d_linear_rowdot_packed":   # hidden [M,N] never written: x + W + one float per row
        M, K, N = d["M"], d["K"], d["N"]
        return 4.0 * (M * K + K * N + N + M), 2.0 * M * K * N + 2.0 * M * N, "tensor"
    if name in ("dh3d_flex_conv_pm", "dh3d_flex_conv_pm_packed"):
        n, K, Ci, Co = d["n"], d["K"], d["Ci"], d["Co"]
        return 4.0 * (n * Ci + n * Co + n * K + 3 * n + 4 * Ci * Co), 8.0 * n * K * Ci + 8.0 * n * Ci * Co, "hbm"
    if name == "dh3d_linear_chain_packed":  # x and both weights once, y once; the hidden [M,H] activation stays on the SM
        M, K, H, N = d["M"], d["K"], d["H"], d["N"]
        return 4.0 * (M * K + K * H + H * N + M * N), 2.0 * M * (K * H + H * N), "hbm"
    if name == "dh3d_linear_join_packed":   # both inputs and weights once, y and its normalised copy once
        M, Ka, Kb, N = d["M"], d["Ka"], d["Kb"], d["N"]
        return 4.0 * (M * (Ka + Kb) + (Ka + Kb) * N + 2 * M * N), 2.0 * M * (Ka + Kb) * N, "hbm"
    if name in ("dh3d_knn_bruteforce_pm", "dh3d_knn_query_sorted"):
        B, N, K = d["B"], d["N"], d["K"]
        return B * (12.0 * N + 8.0 * N * K), 8.0 * B * N * N, "alu"
    if name == "dh3d_netvlad":   # 4*[nD + n + 2*D*Kc + 256] per cloud + the hidden weights once
        B, N, D, Kc, O = d["B"], d["N"], d["D"], d["Kc"], d["O"]
        return 4.0 * (B * (N * D + N + O) + 2 * D * Kc + D * Kc * O), B * 4.0 * N * D * Kc, "hbm"
    if name in ("dh3d_farthest_point_sample", "dh3d_farthest_point_sample_presorted"):
        B, N, M = d["B"], d["N"], d["M"]
        return B * (12.0 * N + 4.0 * M), 8.0 * B * N * (M - 1), "latency"
    if name == "dh3d_flex_pool_pm":          # 4*[2nD + nD(argmax) + nK]
        n, K, D, A = d["n"], d["K"], d["D"], d["A"]
        return 4.0 * (2 * n * D + (n * D if A else 0) + n * K), 0.0, "hbm"
    if name == "dh3d_conv_pointset_pm":      # 4*[3n + 32n + nK]
        n, K, Ci, Co = d["n"], d["K"], d["Ci"], d["Co"]
        return 4.0 * (n * Ci + n * Co + n * K), 2.0 * n * K * Ci * Co, "hbm"
    if name in ("dh3d_group_point", "dh3d_group_point_ld"):   # 4*[m*S*C*2 + m*S]
        B, M, S, C = d["B"], d["M"], d["S"], d["C"]
        return 4.0 * B * (2 * M * S * C + M * S), 0.0, "hbm"
    if name in ("dh3d_three_nn_ws", "dh3d_three_nn_ws_presorted", "dh3d_three_nn_presorted2"):           # 12(n+m) + 24n
        B, n, m = d["B"], d["n"], d["m"]
        return B * (12.0 * (n + m) + 24.0 * n), 8.0 * B * n * m, "alu"
    if name in ("dh3d_three_interpolate", "dh3d_three_interpolate_from_dist", "dh3d_three_interpolate_ld"):
        B, n, m, C = d["B"], d["n"], d["m"], d["C"]
        return 4.0 * B * (m * C + n * C + 6 * n), 6.0 * B * n * C, "hbm"
    if name == "dh3d_se_pool_excite":        # x once, nbr once, out once (+ the pooled gather from L2)
        n, K, C = d["n"], d["K"], d["C"]
        return 4.0 * (2 * n * C + n * K), 4.0 * n * C * (C // 4), "hbm"
    return 0.0, 0.0, None


# C-ABI entry point -> the kernel that dominates it (prefix of the ncu kernel name), for roofline.traffic
OP_KERNEL = {
    "dh3d_linear_rowdot_packed": ("gemm_head16_kernel", "gemm_tc16_kernel"),
    "dh3d_linear_packed": ("gemm_tc16_kernel",),
    "dh3d_netvlad": ("netvlad_tc2_kernel", "netvlad_tc_kernel"),
    "dh3d_knn_bruteforce_pm": ("knn_query_kernel<8, 1, 1",),
    "dh3d_knn_query_sorted": ("knn_query_kernel<8, 1, 1",),
    "dh3d_farthest_point_sample": ("fps_cluster_kernel",),
    "dh3d_farthest_point_sample_presorted": ("fps_bucket_kernel",),
    "dh3d_flex_conv_pm": ("flexconv_ca_kernel",),
    "dh3d_flex_conv_pm_packed": ("flexconv_ca_kernel",),
    "dh3d_linear_join_packed": ("gemm_join16_kernel",),
    "dh3d_linear_chain_packed": ("gemm_chain16_kernel",),
    "dh3d_se_pool_excite": ("se_pool_excite_kernel",),
}


def ncu_traffic(op_name, avg_launch_us):
    """dram read+write bytes per launch of the op's dominant kernel from the committed ncu --set full
    capture (profiles/ncu_traffic.json, made by scripts/ncu_summary.py traffic); the capture whose
    duration is closest to the live launch time is the same shape.  None if there is no capture."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path) or op_name not in OP_KERNEL:
        return None
    with open(path) as f:
        caps = json.load(f)
    cand = [c for c in caps if c["kernel"].startswith(OP_KERNEL[op_name])]
    if not cand:
        return None
    best = min(cand, key=lambda c: abs(c["time_us"] - avg_launch_us))
    if abs(best["time_us"] - avg_launch_us) > 0.5 * avg_launch_us:
        return None
    return best["dram_bytes"]


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU comparator is meant to use every host thread."""
    import oracle
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    oracle.set_num_threads(n)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)   # numpy BLAS for the dense layers
    except Exception:  # noqa: BLE001
        pass
    return n


def workload_name(args):
    if args.workload == "full":
        return ("full DH3D local+detector+global forward, N=8192, batch=%d per GPU "
                "(BASELINE.json configs[2])" % args.batch)
    if args.workload == "local":
        return "local descriptor forward (basic_config), N=8192, batch=%d per GPU (configs[1])" % args.batch
    if args.workload == "retrieval":
        return ("global retrieval: %d synthetic clouds x N=8192 sharded across the GPUs in micro-batches of %d, one "
                "all-gather of the 256-D global descriptors, k=25 retrieval (configs[3])" % (args.total_clouds, args.batch))
    return "FlexConv+kNN sweep N in {4096,8192,16384,32768}, K in {8,16,32}, C=128, B=8 (configs[4])"


def workload_config(args):
    from dh3d_b200.configs import basic_config, full_config, global_config
    if args.workload == "full":
        return full_config()
    if args.workload == "local":
        return basic_config()
    return global_config()


def run_reference_arm(args, rank, world):
    """The reference's CPU path on the host cores: the oracle port (oracle/dh3d_oracle.c literal
    loops incl. the full-sort kNN of knn_bruteforce_kernel.cc, OpenMP over all cores; dense layers
    in numpy) -- one cloud per step (bounded sample)."""
    if rank != 0:
        return
    import oracle
    from oracle import net
    from dh3d_b200.model import DH3D, init_random_
    use_all_host_threads()
    cfg = workload_config(args)
    model = init_random_(DH3D(cfg), seed=0)
    params = {k: v.detach().numpy() for k, v in model.named_parameters()}
    clouds = synth_clouds(1, N_POINTS, 0).numpy()
    steps, warm = max(1, args.steps), min(args.warmup, 1)
    for _ in range(warm):
        net.forward(clouds, params, detection=cfg.detection, extract_global=cfg.extract_global,
                    reference_cpu=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.forward(clouds, params, detection=cfg.detection, extract_global=cfg.extract_global,
                    reference_cpu=True)
    dt = time.perf_counter() - t0
    value = steps * clouds.shape[0] / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "retrieval" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "n_points": N_POINTS, "clouds_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": "1 synthetic cloud of 8192 points per step (of the 32-cloud batch); "
                                   "oracle port of the reference CPU functors, OpenMP over all host threads"},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- small helpers shared by the GPU workloads ------------------------------------------------------------------
class Ctx(object):
    """torch / torch.distributed state of this rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """barrier + sync, `steps` calls of fn(i) between two CUDA events on the current stream, barrier + sync;
        -> max-over-ranks elapsed ms."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def d2h_ceiling(ctx, nbytes, reps=10):
    """Raw pinned-host D2H rate of this box with ALL ranks copying at once (plain cudaMemcpyAsync of the same
    size as a step's result block): the ceiling any end-to-end number that returns that block can reach.
    -> aggregate GB/s over the ranks."""
    torch = ctx.torch
    src = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    ms = ctx.timed(lambda i: dst.copy_(src, non_blocking=True), reps)
    return ctx.world * nbytes * reps / (ms / 1e3) / 1e9


class E2E(object):
    """Host buffers -> public API -> host buffers.  Per step: H2D of that step's clouds from pinned memory, the
    forward (graph replay of the public DH3D.forward), ONE D2H of the step's contiguous result block into pinned
    memory on a copy stream, overlapped with the next step's compute (everything inside the timed region; the
    closing barrier waits for the last copy).  Two graph instances alternate so that a result block is not
    overwritten before its copy has finished."""

    def __init__(self, ctx, graphs, host_batches, post=None):
        torch = ctx.torch
        self.ctx, self.graphs, self.host_batches, self.post = ctx, graphs, host_batches, post
        self.copy_stream = torch.cuda.Stream(device=ctx.dev)
        self.copy_done = [None, None]
        blocks = [g.block if post is None else post.block(k) for k, g in enumerate(graphs)]
        self.blocks = blocks
        self.host_out = [torch.empty(b.shape, dtype=b.dtype).pin_memory() for b in blocks]
        self.h2d = host_batches[0].numel() * host_batches[0].element_size()
        self.d2h = blocks[0].numel() * blocks[0].element_size()

    def step(self, i):
        torch = self.ctx.torch
        k = i % 2
        if self.copy_done[k] is not None:
            torch.cuda.current_stream().wait_event(self.copy_done[k])
        pts = self.host_batches[i % len(self.host_batches)].to(self.ctx.dev, non_blocking=True)
        out = self.graphs[k](pts)
        if self.post is not None:
            self.post(k, pts, out)
        ready = torch.cuda.Event()
        ready.record()
        self.copy_stream.wait_event(ready)
        with torch.cuda.stream(self.copy_stream):
            self.host_out[k].copy_(self.blocks[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self.copy_done[k] = ev

    def run(self, steps, clouds_per_step):
        for i in range(3):
            self.step(i)
        ms = self.ctx.timed(self.step, steps)
        return self.ctx.world * clouds_per_step * steps / (ms / 1e3), ms / steps


class KeypointPost(object):
    """--perform_nms output mode on the device (dh3d_b200.model.extract_keypoints): the step's result block is the
    [B,512,132] keypoint rows + their counts instead of the dense [B,N,132] map."""

    def __init__(self, ctx, B, feat_dim, max_kp=512):
        torch = ctx.torch
        self.rows = [torch.zeros((B * max_kp * (3 + feat_dim + 1) + B,), dtype=torch.float32, device=ctx.dev)
                     for _ in range(2)]
        self.B, self.C, self.max_kp = B, feat_dim, max_kp

    def block(self, k):
        return self.rows[k]

    def __call__(self, k, pts, out):
        from dh3d_b200.model import extract_keypoints
        n = self.B * self.max_kp * (3 + self.C + 1)
        view = self.rows[k][:n].view(self.B, self.max_kp, 3 + self.C + 1)
        _, counts = extract_keypoints(pts, out["local_desc"], out["attention"], max_keypoints=self.max_kp, out=view)
        self.rows[k][n:].copy_(counts.to(self.rows[k].dtype), non_blocking=True)


def op_roofline_table(per_step, peaks):
    """Every timed op of one step against its roof: algorithmic bytes / device ms -> GB/s and fraction of the
    measured HBM peak (for 'hbm' ops that is the grade; 'tensor' ops also get TFLOP/s; 'alu' / 'latency' ops are
    listed with their GB/s for completeness -- their compulsory traffic is tiny)."""
    rows, tot_bytes = [], 0.0
    for key, (ms, calls) in sorted(per_step.items(), key=lambda kv: -kv[1][0]):
        name = key.split("[")[0]
        tag = key.split("[")[1].rstrip("]") if "[" in key else ""
        try:
            byts, flops, bound = algorithmic(tag, name) if tag else (0.0, 0.0, None)
        except KeyError:
            byts, flops, bound = 0.0, 0.0, None
        if bound is None:
            rows.append({"op": key, "ms": round(ms, 4), "calls": calls})
            continue
        byts *= calls
        flops *= calls
        tot_bytes += byts
        gbs = byts / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        row = {"op": key, "ms": round(ms, 4), "calls": calls, "bound": bound, "algorithmic_mb": round(byts / 1e6, 3),
               "gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 4)}
        if bound == "tensor":
            row["tflops"] = round(flops / (ms / 1e3) / 1e12, 1)
            row["frac_bf16_burst"] = round(row["tflops"] / peaks["bf16_tflops"], 4)
        rows.append(row)
    return rows, tot_bytes


def ref_cuda_leg(ctx, B):
    """The reference's OWN CUDA kernels (oracle/_ref/libdh3d_ref_cuda.so: unmodified sources compiled for sm_100a;
    comparator only) against this repo's drop-in entries -- same reference [B,C,N] layouts, same inputs, same GPU,
    same run -- for every custom op of one 32-cloud forward that the reference runs on the GPU (three_nn /
    three_interpolate are CPU-only ops in the reference; NetVLAD and the 1x1 stacks are TensorFlow library ops)."""
    torch = ctx.torch
    try:
        from oracle import ref
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "oracle.ref import failed: %s" % e}
    if not ref.have_cuda():
        return {"unavailable": "oracle/_ref/libdh3d_ref_cuda.so not built (built from /root/reference in the build container)"}
    from dh3d_b200 import ops, tf_ops, user_ops
    g = torch.Generator(device="cuda").manual_seed(0)
    N = N_POINTS

    def timeit(fn, warm=1, reps=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    pts = (torch.rand((B, N, 3), device="cuda", generator=g) * 50 - 25).contiguous()
    pos = pts.transpose(1, 2).contiguous()
    rows = []

    def add(op, shape, ref_fn, mine_fn, native_fn=None, calls=1):
        r, m = timeit(ref_fn), timeit(mine_fn, 2, 5)
        row = {"op": op, "shape": shape, "calls_per_forward": calls, "ref_cuda_ms": round(r, 4),
               "dropin_ms": round(m, 4), "speedup": round(r / m, 2)}
        if native_fn is not None:
            row["native_pm_ms"] = round(timeit(native_fn, 2, 5), 4)
        rows.append(row)

    add("knn_bruteforce", "B%d N8192 K8" % B, lambda: ref.cuda_knn(pos, 8), lambda: user_ops.knn_bruteforce(pos, 8),
        lambda: ops.knn_points(pts, 8))
    nbr, _ = ops.knn_points(pts, 8)
    nc = nbr.transpose(1, 2).contiguous()
    kp = tf_ops.farthest_point_sample(N // 8, pts)
    pts_s = tf_ops.gather_point(pts, kp)
    pos_s = pts_s.transpose(1, 2).contiguous()
    add("knn_bruteforce", "B%d N1024 K8" % B, lambda: ref.cuda_knn(pos_s, 8), lambda: user_ops.knn_bruteforce(pos_s, 8),
        lambda: ops.knn_points(pts_s, 8), calls=2)
    nbr_s, _ = ops.knn_points(pts_s, 8)
    nc_s = nbr_s.transpose(1, 2).contiguous()
    add("farthest_point_sample", "B%d N8192 M1024" % B, lambda: ref.cuda_fps(N // 8, pts),
        lambda: tf_ops.farthest_point_sample(N // 8, pts), calls=2)
    th2 = torch.randn((3, 32), device="cuda", generator=g)
    bi2 = torch.randn((32,), device="cuda", generator=g)
    add("convolution_pointset", "3->32 @%d" % (B * N), lambda: ref.cuda_conv_pointset(pos, nc, th2, bi2),
        lambda: user_ops.convolution_pointset(pos, nc, th2, bi2), lambda: ops.conv_pointset(pts, th2, bi2, nbr))
    for (ci, co, p_cm, n_cm, p_pm, n_pm, npts) in ((32, 64, pos, nc, pts, nbr, N), (64, 64, pos, nc, pts, nbr, N),
                                                   (64, 128, pos_s, nc_s, pts_s, nbr_s, N // 8),
                                                   (128, 128, pos_s, nc_s, pts_s, nbr_s, N // 8),
                                                   (128, 256, pos_s, nc_s, pts_s, nbr_s, N // 8)):
        f = torch.randn((B, npts, ci), device="cuda", generator=g)
        fc = f.transpose(1, 2).contiguous()
        th = torch.randn((3, ci, co), device="cuda", generator=g) / ci ** 0.5
        bi = torch.randn((ci, co), device="cuda", generator=g) / ci ** 0.5
        add("flex_convolution", "%d->%d @%d K8" % (ci, co, B * npts),
            lambda: ref.cuda_flex_conv(fc, p_cm, n_cm, th, bi), lambda: user_ops.flex_convolution(fc, p_cm, n_cm, th, bi),
            lambda: ops.flex_conv(f, th, bi, n_pm, p_pm))
    for (D, n_cm, n_pm, npts) in ((32, nc, nbr, N), (64, nc, nbr, N), (128, nc_s, nbr_s, N // 8)):
        f = torch.randn((B, npts, D), device="cuda", generator=g)
        fc = f.transpose(1, 2).contiguous()
        add("flex_pooling", "D%d @%d K8" % (D, B * npts), lambda: ref.cuda_flex_pool(fc, n_cm),
            lambda: user_ops.flex_pooling(fc, n_cm), lambda: ops.flex_pool(f, n_pm))
    kp3 = kp.unsqueeze(2).contiguous()
    for C, calls in ((64, 1), (128, 1)):
        f = torch.randn((B, N, C), device="cuda", generator=g)
        add("group_point", "C%d M1024 @B%d" % (C, B), lambda: ref.cuda_group_point(f, kp3),
            lambda: tf_ops.group_point(f, kp3), calls=calls)
    tot_r = sum(r["ref_cuda_ms"] * r["calls_per_forward"] for r in rows)
    tot_m = sum(r["dropin_ms"] * r["calls_per_forward"] for r in rows)
    return {"rows": rows, "sum_ref_cuda_ms_per_forward": round(tot_r, 3), "sum_dropin_ms_per_forward": round(tot_m, 3),
            "speedup_custom_ops": round(tot_r / tot_m, 1), "clouds": B,
            "note": "ref_cuda = the reference's unmodified CUDA kernels compiled for sm_100a (oracle/_ref), launched "
                    "through their own launchers on the default stream; dropin = this repo's reference-layout C-ABI "
                    "entries (the ones INTEGRATION.md binds, layout transposes included); native_pm = the point-major "
                    "entries the assembled forward uses.  CUDA events, median of 3-5.  The reference runs three_nn / "
                    "three_interpolate on the CPU and NetVLAD / 1x1 stacks through TensorFlow: not in this table."}


def data_sensitivity_leg(ctx, model, B):
    """The data-dependent kernels (box-pruned k-NN / 3-NN, FPS) and the whole forward on inputs other than the
    friendly U(-25,25)^3: LiDAR-like synthetic, the reference's own demo clouds (tests/golden/demo_clouds.npz, real
    Oxford scans incl. duplicate-padded ones), all-zero padding clouds, far-outlier padding (points at 1e5)."""
    torch = ctx.torch
    import numpy as np
    from dh3d_b200 import ops
    from dh3d_b200.data import synth_lidar_clouds

    def time_ms(fn, reps=8):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    sets = {"uniform": synth_clouds(B, N_POINTS, 0), "lidar_like": torch.from_numpy(synth_lidar_clouds(B, N_POINTS, 0))}
    gold = os.path.join(ROOT, "tests", "golden", "demo_clouds.npz")
    if os.path.exists(gold):
        c = np.load(gold)["clouds"]
        sets["oxford_demo"] = torch.from_numpy(np.concatenate([c] * ((B + len(c) - 1) // len(c)), 0)[:B].copy())
    sets["all_zero"] = torch.zeros((B, N_POINTS, 3))
    far = synth_clouds(B, N_POINTS, 1).clone()
    far[:, N_POINTS - N_POINTS // 8:] = 100000.0
    sets["far_outlier_padding"] = far
    res = {}
    for name, pts in sets.items():
        p = pts.to(ctx.dev)
        m = p[:, :N_POINTS // 8].contiguous()
        fwd = time_ms(lambda: model(p), reps=5)
        res[name] = {"clouds_per_s": round(B / fwd * 1e3, 1), "forward_ms": round(fwd, 4),
                     "knn_ms": round(time_ms(lambda: ops.knn_points(p, 8)), 4),
                     "fps_ms": round(time_ms(lambda: ops.farthest_point_sample(N_POINTS // 8, p)), 4),
                     "three_nn_ms": round(time_ms(lambda: ops.three_nn(p, m)), 4)}
    vals = [v["clouds_per_s"] for k, v in res.items()]
    res["spread_max_over_min"] = round(max(vals) / min(vals), 3)
    res["note"] = ("eager launches (no graph), %d clouds per call; far_outlier_padding = get_fixednum_pcd(randsample="
                   "False) style padding of 1/8 of the points at 1e5 m (core/utils.py:107-108)" % B)
    return res


def cpu_baseline_leg(cfg, B):
    import oracle
    from oracle import net
    from dh3d_b200.model import DH3D, init_random_
    use_all_host_threads()
    cpu_model = init_random_(DH3D(cfg), seed=0)
    params = {k: v.detach().numpy() for k, v in cpu_model.named_parameters()}
    clouds = synth_clouds(B, N_POINTS, 0).numpy()
    done, t0 = 0, time.perf_counter()
    while True:   # whole clouds of the step's batch until ~12 s of CPU work (bounded sample)
        net.forward(clouds[done:done + 1], params, detection=cfg.detection, extract_global=cfg.extract_global,
                    reference_cpu=True)
        done += 1
        dt = time.perf_counter() - t0
        if dt >= 12.0 or done >= B:
            break
    return {"value": done / dt, "unit": "clouds/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d of the %d clouds of one step (N=8192), oracle port of the reference CPU "
                      "functors with OpenMP over all host threads, %.1f s" % (done, B, dt)}


def dominant_roofline(peaks, dominant, tot_ms, calls):
    dom_name = dominant.split("[")[0]
    tag = dominant.split("[")[1].rstrip("]") if "[" in dominant else ""
    abytes, aflops, bound = algorithmic(tag, dom_name) if tag else (0.0, 0.0, None)
    avg_s = (tot_ms / max(calls, 1)) / 1e3
    traffic = ncu_traffic(dom_name, avg_s * 1e6)
    if bound == "tensor":
        ach = aflops / avg_s / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
                "kernel": dominant, "avg_launch_ms": avg_s * 1e3, "launches_timed": calls,
                "peak_source": peaks["source"] + ": bf16 cuBLAS BURST figure (the kernel is timed alone, launch by "
                                                 "launch, in a short eager pass at full clocks)",
                "algorithmic_bytes": abytes,
                "frac_of_sustained_peak": ach / peaks["bf16_tflops_sustained"],
                "note": "achieved = algorithmic 2*M*K*N flops of an fp32-accurate GEMM.  The kernel executes 3 fp16 "
                        "tcgen05 MMAs per product (2-term fp16 split, kind::f16, same rate as bf16), so it issues "
                        "%.0f TFLOP/s of tensor work = %.2f of the burst bf16 peak." % (3 * ach, 3 * ach / peaks["bf16_tflops"])}
    else:
        ach = abytes / avg_s / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "kernel": dominant,
                "avg_launch_ms": avg_s * 1e3, "launches_timed": calls, "peak_source": peaks["source"],
                "algorithmic_bytes": abytes, "algorithmic_flops": aflops}
    if traffic is not None:
        roof["traffic_source"] = "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read+write per launch)"
    return roof


# ---- workloads full / local ---------------------------------------------------------------------------------------
def run_forward_workload(args):
    from dh3d_b200 import _lib
    from dh3d_b200.dist import all_gather_descriptors
    from dh3d_b200.model import DH3D, GraphedForward, InFlightForward, init_random_
    ctx = Ctx(args)
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    cfg = workload_config(args)
    model = init_random_(DH3D(cfg), seed=0).to(dev)
    full = args.workload == "full"
    outputs = ("local_desc", "attention", "globaldesc") if full else ("local_desc",)
    B, K = args.batch, args.steps

    # inputs: R distinct resident batches, rotated (a step's own activations, > 1 GB, sweep the
    # 126 MB L2 between two uses of anything)
    R = 4
    host_batches = [synth_clouds(B, N_POINTS, rank * 1000 + i).pin_memory() for i in range(R)]
    dev_batches = [h.to(dev) for h in host_batches]

    def eager_step(i, pts=None):
        return model(dev_batches[i % R] if pts is None else pts, outputs=outputs)

    graphs, graph_note = None, "eager launches"
    if not args.no_graph:
        try:
            graphs = [GraphedForward(model, dev_batches[0], outputs=outputs) for _ in range(2)]
            graph_note = "forward replayed from a CUDA graph (%d kernels, 2 streams)"
        except Exception as e:  # noqa: BLE001 -- fall back loudly, never silently
            graphs, graph_note = None, "eager launches (graph capture failed: %s)" % str(e)[:120]
            print("bench: CUDA graph capture failed, running eagerly: %s" % e, file=sys.stderr)

    # per-step global descriptors of this rank: kept on the device, gathered ONCE at the end of the timed region
    desc_keep = torch.empty((max(K, 1) * B, cfg.output_dim), dtype=torch.float32, device=dev) if full else None

    def step(i, pts=None):
        out = eager_step(i, pts) if graphs is None else graphs[i % 2](dev_batches[i % R] if pts is None else pts)
        if desc_keep is not None:
            j = i % max(K, 1)
            desc_keep[j * B:(j + 1) * B].copy_(out["globaldesc"], non_blocking=True)
        return out

    # Two steps in flight (model.InFlightForward): graph instance k replays on its own stream ("lane" k), so step i + 1
    # (other instance, other static buffers) starts next to step i and its latency-bound geometry kernels run under
    # step i's tensor-bound heads (2.09 -> 1.94 ms per step, profiles/inflight_r3s.txt; a third lane loses again).
    # Steps stay whole and ordered per lane; the timed region ends when BOTH lanes have drained.
    inflight = InFlightForward(graphs) if graphs is not None and args.in_flight == 2 else None

    def lane_step(i):
        if inflight is None:
            return step(i)
        keep = None
        if desc_keep is not None:
            j = i % max(K, 1)
            keep = lambda out: desc_keep[j * B:(j + 1) * B].copy_(out["globaldesc"], non_blocking=True)  # noqa: E731
        return inflight.submit(dev_batches[i % R], consume=keep)[0]

    def pipelined(n, tail=None):
        """fn(i) for ctx.timed: n steps over the lanes, all lanes joined into the current stream (where ctx.timed
        records its closing event) after the last one, then ``tail``."""
        def fn(i):
            lane_step(i)
            if i == n - 1:
                if inflight is not None:
                    inflight.join()
                if tail is not None:
                    tail()
        return fn

    def gather_tail():
        if desc_keep is not None and world > 1:
            all_gather_descriptors(desc_keep)      # [K*B*world, 256] on every rank: the run's one collective

    warm = pipelined(args.warmup, gather_tail)     # same issue pattern as the timed region; NCCL communicator warm
    for i in range(args.warmup):
        warm(i)
    ctx.barrier()

    # ---- per-op table (untimed pass, all C-ABI calls bracketed by events) -------------------------
    _lib.stats.reset()
    _lib.stats.timing_filter = "all"
    for i in range(2):
        eager_step(i)
    torch.cuda.synchronize()
    op_table = _lib.stats.op_times_ms()
    _lib.stats.timing_filter = None
    per_step = {k: (v[0] / 2.0, v[1] // 2) for k, v in op_table.items()}
    kernels_per_step = _lib.stats.kernels // 2
    dominant = max(per_step.items(), key=lambda kv: kv[1][0])[0]
    dom_name = dominant.split("[")[0]
    if graphs is not None:
        graph_note = graph_note % kernels_per_step

    # ---- host issue time (no sync inside): how long the CPU needs to enqueue one step ---------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        step(i)
    host_issue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, inputs resident, the all-gather at its end ----------------
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    elapsed_ms = ctx.timed(pipelined(K, gather_tail), K)
    clocks = sampler.stop()
    value = world * B * K / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / K
    # the same K steps strictly one after the other on one stream (= the latency of one step)
    serial = None
    if inflight is not None:
        ms1 = ctx.timed(lambda i: step(i), K)
        serial = {"ms_per_step": ms1 / K, "value": world * B * K / (ms1 / 1e3), "unit": "clouds/s",
                  "what": "one step in flight: every replay on one stream (no all-gather); ms_per_step here is a "
                          "step's latency, the headline's is elapsed / steps with two steps in flight"}

    # ---- dominant launch group, timed live with CUDA events in an eager pass over the same steps ---
    _lib.stats.reset()
    _lib.stats.timing_filter = {dom_name}
    for i in range(min(K, 20)):
        eager_step(i)
    torch.cuda.synchronize()
    dom_times = _lib.stats.op_times_ms()
    _lib.stats.timing_filter = None

    # ---- sustained: >= 2 s of back-to-back steps with its own clock / power record ----------------
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(K, int(math.ceil(args.sustained_s * 1e3 / ms_per_step)))
        s2 = ClockSampler(ctx.local_rank)
        s2.start()
        ms = ctx.timed(pipelined(n_sus), n_sus)
        c2 = s2.stop()
        sustained = {"seconds": ms / 1e3, "steps": n_sus, "value": world * B * n_sus / (ms / 1e3), "unit": "clouds/s",
                     "ms_per_step": ms / n_sus, "clocks": c2}

    # ---- e2e: host buffers -> public API -> host buffers ------------------------------------------
    if graphs is None:
        raise SystemExit("bench: the e2e leg replays the public forward from CUDA graphs; --no-graph skips it")
    e2e_run = E2E(ctx, graphs, host_batches)
    e2e_value, e2e_ms = e2e_run.run(K, B)
    ceil_gbs = d2h_ceiling(ctx, e2e_run.d2h)
    ceil_clouds = ceil_gbs * 1e9 / (e2e_run.d2h / B)
    e2e = {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": e2e_run.h2d, "d2h_bytes_per_step": e2e_run.d2h,
           "ms_per_step": e2e_ms, "d2h_copies_per_step": 1,
           "d2h_achieved_gbs": world * e2e_run.d2h / (e2e_ms / 1e3) / 1e9,
           "d2h_ceiling_gbs": ceil_gbs,
           "d2h_ceiling_clouds_per_s": ceil_clouds,
           "frac_of_ceiling": e2e_value / min(ceil_clouds, value),
           "note": "outputs = %s of every cloud (the dense maps the reference's --save_all mode writes); the ceiling is "
                   "plain pinned cudaMemcpyAsync D2H of the same block size with all %d rank(s) copying at once, so "
                   "e2e <= min(device-timed value, ceiling)" % ("+".join(outputs), world)}

    # ---- the reference's other output modes (their D2H is KBs..MBs per step instead of 135 MB) ----
    e2e_modes = None
    if full and not args.no_modes:
        e2e_modes = {}
        g_glob = [GraphedForward(model, dev_batches[0], outputs=("globaldesc",)) for _ in range(2)]
        run = E2E(ctx, g_glob, host_batches)
        v, ms = run.run(K, B)
        dev_ms = ctx.timed(lambda i: g_glob[i % 2](dev_batches[i % R]), K) / K
        e2e_modes["globaldesc_only"] = {
            "value": v, "unit": "clouds/s", "ms_per_step": ms, "device_value": world * B / (dev_ms / 1e3),
            "h2d_bytes_per_step": run.h2d, "d2h_bytes_per_step": run.d2h,
            "what": "outputs=('globaldesc',): what evaluate/global_eval/globaldesc_extract.py fetches (the detector "
                    "head is not needed for it and is not run, as in the reference's pruned TF graph)"}
        del g_glob, run
        post = KeypointPost(ctx, B, cfg.featdim)
        run = E2E(ctx, graphs, host_batches, post=post)
        v, ms = run.run(max(3, K // 4), B)
        e2e_modes["nms_keypoints"] = {
            "value": v, "unit": "clouds/s", "ms_per_step": ms, "h2d_bytes_per_step": run.h2d,
            "d2h_bytes_per_step": run.d2h, "steps": max(3, K // 4),
            "what": "--perform_nms mode of evaluate/local_eval/localdesc_extract.py:92-104 on the device: full forward, "
                    "keypoint NMS on 1 - attention (K = 50 neighbour lists, fp64 radius tests), only the <= 512 detected "
                    "rows of xyz_feat_att per cloud return"}
        del run, post

    ctx.barrier()
    if rank != 0:
        ctx.close()
        return

    # ---- roofline of the dominant launch group + every op against the HBM roof --------------------
    peaks = load_peaks()
    tot_ms, calls = dom_times[dominant] if dominant in dom_times else per_step[dominant]
    roof = dominant_roofline(peaks, dominant, tot_ms, calls)
    op_rows, step_bytes = op_roofline_table(per_step, peaks)
    io_bytes = B * (N_POINTS * 3 * 4) + e2e_run.d2h
    step_hbm = {"algorithmic_bytes_per_step": step_bytes, "algorithmic_mb_per_cloud": step_bytes / B / 1e6,
                "gbs": step_bytes / (ms_per_step / 1e3) / 1e9,
                "frac_hbm": step_bytes / (ms_per_step / 1e3) / 1e9 / peaks["hbm_gbs"],
                "compulsory_io_mb_per_cloud": io_bytes / B / 1e6,
                "compulsory_io_frac_hbm": io_bytes / (ms_per_step / 1e3) / 1e9 / peaks["hbm_gbs"],
                "note": "sum over the step's ops of their per-op compulsory bytes (SURVEY 8d: nothing fused across ops) "
                        "/ device-timed step; the step also runs ~10.6 GFLOP/cloud of dense 1x1 layers, so its roofline "
                        "is max(bytes/BW_HBM, dense_flops/peak): see op_roofline for which op sits on which roof"}

    cpu = ref_cuda = sens = None
    if world == 1:
        if not args.no_ref_cuda:
            ref_cuda = ref_cuda_leg(ctx, B)
        if not args.no_sensitivity and full:
            sens = data_sensitivity_leg(ctx, model, B)
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_leg(cfg, B)

    line = {
        "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "n_points": N_POINTS, "clouds_per_gpu_per_step": B,
                   "knn": 8, "sampled_points": N_POINTS // 8, "weights": "random init (seed 0)",
                   "parallelism": ("clouds sharded by rank, no data-path collective; ONE all_gather of the ranks' "
                                   "[steps*B,256] global descriptors at the end of the timed region")
                   if world > 1 else "single GPU",
                   "launch": graph_note + ("; two steps in flight (dh3d_b200.model.InFlightForward: graph instance k on "
                                           "its own stream)" if inflight is not None else ""),
                   "steps_in_flight": 2 if inflight is not None else 1,
                   "l2": "inputs rotate over %d resident batches; one step streams > 1 GB of activations "
                         "through the 126 MB L2, so nothing survives between steps" % R},
        "e2e": e2e,
        "e2e_modes": e2e_modes,
        "gpu_launches": kernels_per_step * K,
        "host_issue_ms_per_step": host_issue_ms,
        "clocks": clocks,
        "roofline": roof,
        "step_hbm": step_hbm,
        "op_roofline": op_rows,
        "sustained": sustained,
        "one_step_in_flight": serial,
        "cpu_baseline": cpu,
        "ref_cuda": ref_cuda,
        "data_sensitivity": sens,
    }
    if args.op_table:
        with open(args.op_table, "w") as f:
            json.dump({k: {"ms_per_step": v[0], "calls_per_step": v[1]} for k, v in per_step.items()}, f, indent=1)
    print(json.dumps(line), flush=True)
    ctx.close()


# ---- workload retrieval (configs[3]) ------------------------------------------------------------------------------
def run_retrieval_workload(args):
    from dh3d_b200 import _lib
    from dh3d_b200.dist import all_gather_descriptors, shard_range
    from dh3d_b200.model import DH3D, GraphedForward, InFlightForward, init_random_
    from dh3d_b200.retrieval import retrieve_topk
    ctx = Ctx(args)
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    cfg = workload_config(args)
    model = init_random_(DH3D(cfg), seed=0).to(dev)
    B = args.batch
    total = args.total_clouds
    lo, hi = shard_range(total, rank, world)
    n_local = hi - lo
    if n_local % B or total % world:
        raise SystemExit("retrieval: %d clouds do not split into whole micro-batches of %d on %d ranks" % (total, B, world))
    K = n_local // B                       # micro-batches of this rank = steps
    # every cloud of the job is distinct (seeded by its global id), resident copies for the device-timed leg
    host_all = torch.cat([synth_clouds(1, N_POINTS, 100000 + c) for c in range(lo, hi)], 0).pin_memory()
    dev_all = host_all.to(dev)
    host_batches = [host_all[i * B:(i + 1) * B] for i in range(K)]
    dev_batches = [dev_all[i * B:(i + 1) * B] for i in range(K)]
    R = K
    _lib.stats.reset()
    model(dev_batches[0], outputs=("globaldesc",))          # one eager forward: the kernels a graph replay launches
    kernels_per_replay = _lib.stats.kernels
    graphs = [GraphedForward(model, dev_batches[0], outputs=("globaldesc",)) for _ in range(2)]
    desc = torch.empty((n_local, cfg.output_dim), dtype=torch.float32, device=dev)
    topk_host = torch.empty((n_local, 25), dtype=torch.int32).pin_memory()
    state = {}

    def finish():
        all_desc = all_gather_descriptors(desc)                 # [total, 256] on every rank (4 MiB at 4096 clouds)
        idx, _ = retrieve_topk(all_desc, desc, 25)               # this rank's clouds as queries against all of them
        state["idx"] = idx
        return idx

    # two micro-batches in flight (model.InFlightForward: graph instance k replays on its own stream); the job's tail
    # (all-gather + retrieval) runs on the current stream once both lanes have drained
    inflight = InFlightForward(graphs) if args.in_flight == 2 else None

    def run_batch(i, pts):
        def keep(out):
            desc[i * B:(i + 1) * B].copy_(out["globaldesc"], non_blocking=True)
        if inflight is None:
            keep(graphs[i % 2](pts))
        else:
            inflight.submit(pts, consume=keep)

    def laned(source, tail):
        def fn(i):
            run_batch(i, source(i))
            if i == K - 1:
                if inflight is not None:
                    inflight.join()
                tail()
        return fn

    step_resident = laned(lambda i: dev_batches[i % R], finish)
    # end to end: H2D of each micro-batch from pinned memory (current stream; the lane waits for it), and the job's
    # result, [n_local, 25] neighbour ids, back to the host
    step_e2e = laned(lambda i: host_batches[i % R].to(dev, non_blocking=True),
                     lambda: topk_host.copy_(finish(), non_blocking=True))

    for i in range(max(args.warmup, 3)):
        graphs[i % 2](dev_batches[i % R])
    finish()
    ctx.barrier()
    _lib.stats.reset()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ms = ctx.timed(step_resident, K)
    clocks = sampler.stop()
    kernels = _lib.stats.kernels + kernels_per_replay * K    # eager C-ABI calls (top-k) + K graph replays
    value = total / (ms / 1e3)
    ms_e2e = ctx.timed(step_e2e, K)
    e2e_value = total / (ms_e2e / 1e3)
    # the all-gather + retrieval tail alone
    ms_tail = ctx.timed(lambda i: finish(), 5) / 5
    ctx.barrier()
    if rank != 0:
        ctx.close()
        return
    idx = state["idx"].cpu()
    self_first = float((idx[:, 0].long() == torch.arange(lo, hi)).float().mean())   # a cloud's nearest descriptor is its own
    line = {
        "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "n_points": N_POINTS, "total_clouds": total,
                   "clouds_per_gpu": n_local, "micro_batch": B, "outputs": "globaldesc only",
                   "parallelism": "contiguous shards of clouds per rank, no data-path collective; one all_gather_into_tensor "
                                  "of [%d,256] per rank (%d KiB) then dh3d_topk_l2 k=25" % (n_local, n_local),
                   "launch": "forward replayed from a CUDA graph" + ("; two micro-batches in flight (dh3d_b200.model."
                                                                      "InFlightForward)" if inflight is not None else ""),
                   "steps_in_flight": 2 if inflight is not None else 1, "retrieval_self_match_first": self_first},
        "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": B * N_POINTS * 12,
                "d2h_bytes_per_step": n_local * 25 * 4 / K, "ms_total": ms_e2e,
                "note": "per micro-batch: H2D of 32 clouds from pinned memory + forward; at the end one all-gather of the "
                        "descriptors, the k=25 retrieval and ONE D2H of this rank's [%d,25] neighbour ids" % n_local},
        "gather_plus_retrieval_ms": ms_tail,
        "gpu_launches": kernels, "clocks": clocks, "ms_total": ms,
        "roofline": None, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)
    ctx.close()


# ---- workload sweep (configs[4]) ----------------------------------------------------------------------------------
def run_sweep_workload(args):
    from dh3d_b200 import ops, user_ops
    ctx = Ctx(args)
    if ctx.world > 1:
        if ctx.rank == 0:
            print(json.dumps({"metric": "flexconv_knn_sweep", "unavailable": "the sweep is a 1-GPU workload (configs[4])"}))
        ctx.close()
        return
    torch = ctx.torch
    peaks = load_peaks()
    hbm = peaks["hbm_gbs"]
    g = torch.Generator(device="cuda").manual_seed(0)
    B, C = 8, 128
    reps = max(args.steps, 5) if args.steps else 10

    def time_ms(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    rows = []
    for N in (4096, 8192, 16384, 32768):
        pts = (torch.rand((B, N, 3), device="cuda", generator=g) * 50 - 25).contiguous()
        pos = pts.transpose(1, 2).contiguous()
        f = torch.randn((B, N, C), device="cuda", generator=g)
        fc = f.transpose(1, 2).contiguous()
        th = torch.randn((3, C, C), device="cuda", generator=g) / C ** 0.5
        bi = torch.randn((C, C), device="cuda", generator=g) / C ** 0.5
        packed = ops.flex_conv_prepack(th, bi)
        for K in (8, 16, 32):
            knn_ms = time_ms(lambda: ops.knn_points(pts, K))
            nbr, _ = ops.knn_points(pts, K)
            nc = nbr.transpose(1, 2).contiguous()
            fc_ms = time_ms(lambda: ops.flex_conv_packed(f, packed, nbr, pts))
            cm_ms = time_ms(lambda: user_ops.flex_convolution(fc, pos, nc, th, bi))
            n = B * N
            fbytes = 4.0 * (n * C + n * C + n * K + 3 * n + 4 * C * C)
            kbytes = B * (12.0 * N + 8.0 * N * K)
            rows.append({"N": N, "K": K, "B": B, "C": C,
                         "flexconv_ms": round(fc_ms, 4), "flexconv_gbs": round(fbytes / fc_ms / 1e6, 1),
                         "flexconv_frac_hbm": round(fbytes / fc_ms / 1e6 / hbm, 4),
                         "flexconv_gathered_l2_gbs": round(4.0 * n * K * C / fc_ms / 1e6, 1),
                         "flexconv_dropin_cm_ms": round(cm_ms, 4),
                         "flexconv_dropin_frac_hbm": round(fbytes / cm_ms / 1e6 / hbm, 4),
                         "knn_ms": round(knn_ms, 4), "knn_gbs": round(kbytes / knn_ms / 1e6, 1),
                         "knn_gpairs_per_s": round(B * float(N) * N / knn_ms / 1e6, 1)})
    fr = [r["flexconv_frac_hbm"] for r in rows]
    line = {"metric": "flexconv_knn_sweep_algorithmic_hbm_gbs", "value": sum(r["flexconv_gbs"] for r in rows) / len(rows),
            "unit": "GB/s", "n_gpus": 1, "steps": reps, "warmup": max(args.warmup, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "timing": "CUDA events per call, median of %d; every call's "
                       "working set (>= 34 MB of gathered rows, up to 0.5 GB) streams through L2" % reps},
            "hbm_peak_gbs": hbm, "flexconv_frac_hbm_min": min(fr), "flexconv_frac_hbm_max": max(fr), "sweep": rows,
            "note": "value = mean FlexConv algorithmic GB/s over the grid; algorithmic bytes = SURVEY 8(d) "
                    "4*[n*Ci + n*Co + n*K + 3n + 4*Ci*Co]; k-NN is ALU-bound (pairs/s), its GB/s is listed as asked"}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dh3d_b200", choices=["dh3d_b200", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "local", "retrieval", "sweep"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--total-clouds", type=int, default=4096, help="retrieval workload: clouds over all ranks")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="length of the sustained loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-sensitivity", action="store_true")
    ap.add_argument("--no-modes", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--in-flight", type=int, default=2, choices=[1, 2],
                    help="steps in flight on the device (2 = the two graph instances replay on their own streams)")
    ap.add_argument("--op-table", default=None, help="write the per-op device-time table (JSON) here")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 8 if args.workload == "local" else 32
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        run_reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    args.warmup = max(args.warmup, 3)
    if args.workload in ("full", "local"):
        args.steps = 50 if args.steps is None else max(1, args.steps)
        run_forward_workload(args)
    elif args.workload == "retrieval":
        run_retrieval_workload(args)
    else:
        run_sweep_workload(args)


if __name__ == "__main__":
    main()
