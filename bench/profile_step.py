"""Kernel-time breakdown of one fused PS round (torch.profiler; run under gpurun)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/profile_step.txt")
    ap.add_argument("--no-direct-grads", action="store_true")
    ap.add_argument("--no-overlap-wgrad", action="store_true")
    a = ap.parse_args()
    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    from byzpy_b200.attacks import SignFlipAttack
    from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode
    from byzpy_b200.engine.parameter_server.ps import ParameterServer
    from byzpy_b200.models import build_model

    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    xs, ys = B.make_pool(a.workers, a.batch, 224, 1000, 2, 0)
    hon, byz = [], []
    for g in range(a.workers):
        torch.manual_seed(0)
        m = build_model("resnet18", num_classes=1000)
        kw = dict(lr=0.05, momentum=0.9, device=str(dev), preprocess=B.preprocess_fused)
        if g < a.workers - 2:
            hon.append(DeviceHonestNode(m, **kw))
        else:
            byz.append(DeviceByzantineNode(SignFlipAttack(), model=m, **kw))
    ps = ParameterServer(hon, byz, CoordinateWiseMedian(), update_byzantines=True, fused=True,
                         use_cuda_graph=False, worker_streams=a.streams,
                         direct_grads=not a.no_direct_grads, overlap_wgrad=not a.no_overlap_wgrad)
    bt = [(xs[s][0], ys[s][0]) for s in range(a.workers)]
    for _ in range(3):
        ps.step(bt)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            ps.step(bt)
        torch.cuda.synchronize()
    txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
    with open(a.out, "w") as f:
        f.write(txt)
    print(txt[-3000:])
    # stream occupancy: how much of the wall span do kernels cover (gaps = launch / dependency latency)
    ks = []
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower():
            ks.append((float(e.time_range.start), float(e.time_range.end), e.name))
    if ks:
        ks.sort()
        span = ks[-1][1] - ks[0][0]
        busy = sum(k[1] - k[0] for k in ks)
        # union of intervals (kernels on different streams overlap)
        cover, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
        for s_, e_, _ in ks[1:]:
            if s_ > cur_e:
                cover += cur_e - cur_s
                cur_s, cur_e = s_, e_
            else:
                cur_e = max(cur_e, e_)
        cover += cur_e - cur_s
        line = (f"device activities: {len(ks)}  span {span / 1e3:.3f} ms  sum of durations {busy / 1e3:.3f} ms  "
                f"covered {cover / 1e3:.3f} ms  idle {100 * (1 - cover / span):.1f} %")
        print(line)
        with open(a.out, "a") as f:
            f.write("\n" + line + "\n")


if __name__ == "__main__":
    main()
