"""Experiment (B200): does keeping TWO (or three) steps in flight -- CUDA-graph instances of the full forward replayed on
their own streams -- raise whole-job throughput?  Prints clouds/s for serial replays on one stream / 2 lanes / 3 lanes.
(profiles/inflight_priority_r3r.txt also holds the stream-priority variant of this run: capturing the main chain on a
higher-priority stream than the geometry side stream LOSES, 2.07 -> 2.13 ms serial, so GraphedForward has no such knob.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_clouds, workload_config, N_POINTS
from dh3d_b200.model import DH3D, GraphedForward, init_random_


class A(object):
    workload, batch = "full", 32


def main():
    dev = torch.device("cuda:0")
    cfg = workload_config(A)
    model = init_random_(DH3D(cfg), seed=0).to(dev)
    B, K, R = 32, 40, 4
    batches = [synth_clouds(B, N_POINTS, i).to(dev) for i in range(R)]
    outputs = ("local_desc", "attention", "globaldesc")

    def run(graphs, streams, label):
        cur = torch.cuda.current_stream()
        def loop(n):
            for i in range(n):
                s = streams[i % len(streams)]
                if s is None:
                    graphs[i % len(graphs)](batches[i % R])
                else:
                    with torch.cuda.stream(s):
                        graphs[i % len(graphs)](batches[i % R])
        loop(6)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            for s in streams:
                if s is not None:
                    s.wait_event(e0)
            loop(K)
            for s in streams:
                if s is not None:
                    cur.wait_stream(s)
            e1.record(cur)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("%-58s %.4f ms/step  %.0f clouds/s" % (label, best / K, B * K / best * 1e3), flush=True)

    graphs = [GraphedForward(model, batches[0], outputs=outputs) for _ in range(3)]
    run(graphs[:2], [None], "serial, one stream")
    s2 = [torch.cuda.Stream() for _ in range(3)]
    run(graphs[:2], s2[:2], "2 steps in flight (2 streams)")
    run(graphs[:3], s2[:3], "3 steps in flight (3 streams)")
    # results of overlapped replays = results of serial replays, bit for bit
    want = {k: v.clone() for k, v in graphs[0](batches[1]).items()}
    torch.cuda.synchronize()
    for rep in range(4):
        for k in range(3):
            with torch.cuda.stream(s2[k]):
                graphs[k](batches[1])
    torch.cuda.synchronize()
    for g in graphs:
        for k, v in want.items():
            assert torch.equal(g.static_out[k], v), k
    print("overlapped replays are bit-identical to serial ones")


if __name__ == "__main__":
    main()
