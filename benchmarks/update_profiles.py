#!/usr/bin/env python
"""Regenerates profiles/ (tracked, judged) from the scratch artefacts a `scratch/final_run.sh` gpurun call left in
gpurun_out/: the bench lines, the ncu launch list and the `--set full` capture of the fused kernel.

    python benchmarks/update_profiles.py [--round r01]
"""
import argparse
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "memory_l1_wavefronts_shared_ideal", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "sm__icc_request_hit_rate.pct", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
)


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--clips", type=int, default=512)
    a = ap.parse_args()
    r = a.round
    os.makedirs(P, exist_ok=True)
    for src, dst in ((f"bench_{r}_final.json", f"{r}_bench_n1.json"), (f"bench_{r}_reference.json", f"{r}_bench_reference_arm.json"),
                     (f"configs_{r}.json", f"{r}_configs_all.json"), (f"launches_{r}.csv", f"{r}_launches_raw.csv")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    # launch-list summary: time per kernel name and the fused kernel's share of the library's own launches
    lp = os.path.join(G, f"launches_{r}.csv")
    if os.path.exists(lp):
        rows = [x for x in csv.reader(open(lp)) if len(x) > 10]
        h = rows[0]
        ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg, cnt = collections.Counter(), collections.Counter()
        for x in rows[1:]:
            v = float(x[iv].replace(",", ""))
            v = v / 1e3 if x[iu] in ("ns", "nsecond") else (v * 1e3 if x[iu] in ("ms", "msecond") else v)  # -> us
            agg[x[ik]] += v
            cnt[x[ik]] += 1
        own = {k: v for k, v in agg.items() if "b2a::" in k}
        fused = sum(v for k, v in own.items() if "fast_logmel" in k)
        with open(os.path.join(P, f"{r}_launches_summary.csv"), "w") as f:
            f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of\n")
            f.write(f"#   python bench.py --clips {a.clips} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e   (B200)\n")
            f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
            f.write("# torch kernels are the synthetic-input generation outside the timed region.\n")
            f.write(f"# share of the fused kernel among the library's own launches: {fused / max(sum(own.values()), 1e-9):.4f}\n")
            f.write("kernel,launches,total_us,mean_us\n")
            for k, v in agg.most_common():
                f.write(f"\"{k[:140]}\",{cnt[k]},{v:.1f},{v / cnt[k]:.1f}\n")
    rep = os.path.join(G, f"prof_{r}_fast_logmel.ncu-rep")
    if os.path.exists(rep):
        raw = ncu_raw(rep)
        out = {}
        for k, (v, u) in raw.items():
            if k in KEEP or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                out[k] = {"value": v, "unit": u}
        out["_kernel"] = raw.get("Kernel Name", ("", ""))[0]
        out["_command"] = (f"ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 2 -c 1 "
                           f"python bench.py --clips {a.clips} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e")
        tiles = a.clips * 94
        out["_note"] = f"{a.clips} clips x 94 tiles = {tiles} tiles of 32 frames per launch"
        json.dump(out, open(os.path.join(P, f"{r}_fast_logmel_400x160_ncu_full.json"), "w"), indent=1)

        def num(name):
            v, u = raw[name]
            x = float(v.replace(",", ""))
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

        dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        json.dump({"fast_logmel_400x160": {
            "dram_bytes_per_clip": dram / a.clips, "captured_clips": a.clips, "dram_bytes_per_launch_captured": dram,
            "algorithmic_bytes_per_clip": 3456000, "source": f"profiles/{r}_fast_logmel_400x160_ncu_full.json"}},
            open(os.path.join(P, "traffic.json"), "w"), indent=1)
    # the other kernels of the path: same summary per capture
    for tag, what in (("fast_logmel_400_mt", "Hugging Face extractor of Qwen3-ASR, 1024 x 30 s -> (B, 128, 3000): generated mel hf_whisper128, (M, T) "
                       "write-out (benchmarks/bench_configs.py --only V)"),
                      ("fast_logmel_512", "C3 Parakeet 16 x 1 h (benchmarks/bench_configs.py --only C3)"),
                      ("fast_logmel_1024", "C5 Vocos mel forward B=8192 (benchmarks/bench_configs.py --only C5)"),
                      ("fast_istft_1024", "C5 Vocos iSTFT head B=1024 (benchmarks/bench_configs.py --only C5)"),
                      ("istft_small", "C4 Kokoro iSTFT B=1024 (benchmarks/bench_configs.py --only C4)"),
                      ("fast_stft_400", "dsp.stft 400/160, 1024 x 30 s, complex64 rows (benchmarks/bench_configs.py --only S)"),
                      ("resample", "load_audio: 1 h of 44.1 kHz stereo int16 -> 16 kHz mono (benchmarks/bench_configs.py --only R)")):
        rep = os.path.join(G, f"prof_{r}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = ncu_raw(rep)
        out = {k: {"value": v, "unit": u} for k, (v, u) in raw.items()
               if k in KEEP or ("issue_stalled" in k and k.endswith("per_issue_active.ratio"))}
        out["_kernel"] = raw.get("Kernel Name", ("", ""))[0]
        out["_workload"] = what
        out["_command"] = "ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 (scratch/final_run.sh)"
        json.dump(out, open(os.path.join(P, f"{r}_{tag}_ncu_full.json"), "w"), indent=1)
    print("profiles/ updated:", sorted(os.listdir(P)))


if __name__ == "__main__":
    main()
