"""Developer tool: turn the chrome trace `bench.py --trace` writes (rank 0, a few row-sharded steps) into a per-step table.
    python tools/trace_summary.py gpurun_out/g8_trace.json.gz [--step 2] > profiles/<name>.md
One steady-state step = from one owner-compute forward kernel's end to the next one's end.  For each device activity of that
window: stream, start relative to the window, duration.  Then the accounting the scheduling decisions rest on: time covered by
this repo's kernels, by NCCL kernels alone (no kernel of ours running: exposed exchange), and by nothing."""
import argparse
import gzip
import json


def short(name):
    for k, v in (("pair_fwd_kernel", "pair_fwd_kernel<PARTIAL>"), ("bin_scatter_kernel", "bin_scatter_kernel"),
                 ("shard_prep_bins_kernel<1>", "shard_prep_bins_kernel<regen uniform> (PREP_NEG, look-ahead)"),
                 ("shard_prep_bins_kernel<2>", "shard_prep_bins_kernel<regen popular> (PREP_NEG, look-ahead)"),
                 ("shard_prep_bins_kernel<0>", "shard_prep_bins_kernel<ids> (PREP_POS)"), ("shard_finish_kernel", "shard_finish_kernel"),
                 ("bin_scan_kernel", "bin_scan_kernel"), ("loss_sum_kernel", "loss_sum_kernel"), ("gather_rows_kernel", "gather_rows_kernel"),
                 ("AllGather", "NCCL all-gather"), ("AllReduce", "NCCL all-reduce"), ("elementwise_kernel", "torch elementwise (state += inc / id cast)"),
                 ("Memset", "memset"), ("Memcpy", "memcpy DtoD")):
        if k in name:
            return v
    return name[:60]


def union(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def length(iv):
    return sum(b - a for a, b in iv)


def subtract(iv, cut):
    """parts of the (merged) intervals iv not covered by the (merged) intervals cut"""
    out = []
    for a, b in iv:
        cur = a
        for c, d in cut:
            if d <= cur or c >= b:
                continue
            if c > cur:
                out.append([cur, c])
            cur = max(cur, d)
        if cur < b:
            out.append([cur, b])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("trace")
    ap.add_argument("--step", type=int, default=2, help="which steady-state window to print (0 = first)")
    a = ap.parse_args()
    op = gzip.open if a.trace.endswith(".gz") else open
    with op(a.trace, "rt") as f:
        d = json.load(f)
    ev = [e for e in d["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    fwd_end = [e["ts"] + e["dur"] for e in ev if "pair_fwd_kernel" in e["name"]]
    if len(fwd_end) < a.step + 2:
        raise SystemExit("not enough steps in the trace")
    w0, w1 = fwd_end[a.step], fwd_end[a.step + 1]
    win = [e for e in ev if e["ts"] + e["dur"] > w0 + 1e-6 and e["ts"] < w1 - 1e-6]      # everything overlapping the window
    periods = [b - a_ for a_, b in zip(fwd_end, fwd_end[1:])]
    print("# device timeline of one row-sharded step (rank 0; CUPTI through torch.profiler, `bench.py --trace`)\n")
    print("window = end of one owner-compute forward kernel to the end of the next; step periods in this trace: "
          + ", ".join("%.0f" % p for p in periods) + " us\n")
    print("| start us | dur us | stream | activity |")
    print("|---:|---:|---:|---|")
    for e in win:
        print("| %.1f | %.1f | %s | %s |" % (e["ts"] - w0, e["dur"], e["args"].get("stream"), short(e["name"])))
    ours = union([[max(e["ts"], w0), min(e["ts"] + e["dur"], w1)] for e in win if "rsb::" in e["name"]])
    nccl = union([[max(e["ts"], w0), min(e["ts"] + e["dur"], w1)] for e in win if "nccl" in e["name"]])
    anyk = union([[max(e["ts"], w0), min(e["ts"] + e["dur"], w1)] for e in win])
    exposed = subtract(nccl, ours)
    total = w1 - w0
    print("\n| accounting of the window | us | share |")
    print("|---|---:|---:|")
    for label, v in (("window", total), ("a kernel of this repo running", length(ours)), ("NCCL kernel running, none of ours (exposed exchange)", length(exposed)),
                     ("nothing running (launch gaps)", total - length(anyk))):
        print("| %s | %.0f | %.3f |" % (label, v, v / total))
    byname = {}
    for e in win:
        byname.setdefault(short(e["name"]), []).append(e["dur"])
    print("\n| activity | launches | total us |")
    print("|---|---:|---:|")
    for k, v in sorted(byname.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f |" % (k, len(v), sum(v)))


if __name__ == "__main__":
    main()
