#!/usr/bin/env python
"""profiles/r02_summary.md from the bench JSON lines and ncu summaries kept in profiles/ (no GPU needed).
(tools/make_summary_r01.py is the round-1 generator.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def L(name):
    return json.load(open(os.path.join(P, name)))


ns = [n for n in (1, 2, 4, 8) if os.path.exists(os.path.join(P, f"r02_bench_cfg2_n{n}.json"))]
lines = {n: L(f"r02_bench_cfg2_n{n}.json") for n in ns}
n1, ref = lines[1], L("r02_bench_reference_cpu.json")
r1 = L("r01_bench_cfg2_n1.json")
out = ["# Round 2 -- measurements in one place (B200, sm_100a; every number below is in a JSON / txt file of this directory)\n",
       "## Headline (cfg2: DeiT-Small / CIFAR-100 shape, batch 256 per GPU, K = 100 class schemas, class side recomputed every step)\n",
       "| GPUs | images/s (device-resident) | ms/step | images/s end to end (pinned host buffers in, logits out) | H2D GB/s per rank (with kernels / copies only) | file |",
       "|---|---|---|---|---|---|"]
for n in ns:
    d = lines[n]
    e = d["e2e"]
    out.append(f"| {n} | {d['value']:,.0f} ({d['value'] / n1['value']:.2f}x) | {d['ms_per_step']:.3f} | {e['value']:,.0f} | "
               f"{e.get('h2d_GBps_per_rank', 0):.1f} / {e.get('h2d_only_GBps_per_rank', 0):.1f} | r02_bench_cfg2_n{n}.json |")
out.append(f"\nRound 1 at N = 1: {r1['value']:,.0f} images/s ({r1['ms_per_step']:.3f} ms).  The end-to-end arm is bound by the host: "
           "one rank saturates its PCIe link (54-55 GB/s); at N = 8 the copies-only rate is 23 GB/s per rank, i.e. the host's "
           "memory / root-complex ceiling, and the arm with kernels reaches 99 % of it.\n")
out.append(f"Reference head on the box's host CPU ({n1['cpu_baseline']['cores']} cores; the reference's own C++ loops compiled in place + "
           f"its Python glue restated on the same ATen ops): {n1['cpu_baseline']['value']:.0f} images/s (same run) / {ref['value']:.0f} images/s "
           f"(`bench.py --impl reference`, {ref['steps']} timed steps of {ref['ms_per_step']:.0f} ms, r02_bench_reference_cpu.json).  "
           f"Clocks during the timed region: {n1['clocks']['sm_mhz']:.0f} / {n1['clocks']['sm_max_mhz']:.0f} MHz, throttle reasons "
           f"{n1['clocks']['reasons']}.  In-bench parity of the timed path against that CPU head: max rel logit error "
           f"{n1['parity']['logits_max_rel']:.2e} (bar 1e-5), codeword indices equal: {n1['parity']['codeword_indices_equal']}.\n")
out += ["## BASELINE's larger configs (sub-lines `configs.cfg3` / `configs.cfg4` of the same JSON lines)\n",
        "cfg3 = DeiT-Base / Caltech-101 shape, global batch 512 (strong scaling: the batch is sharded); cfg4 = DeiT-Base / ImageNet-1k shape, "
        "1024 images per GPU, K = 1000, D = 1024.  For N > 1 the K class schemas are sharded over the ranks and feat_class is all-gathered by "
        "NCCL inside the captured CUDA graph.\n",
        "| GPUs | cfg3 images/s | cfg3 ms/step | cfg4 images/s | cfg4 ms/step | all-gather (cfg3 / cfg4) | gathered rows vs local recompute |",
        "|---|---|---|---|---|---|---|"]
for n in ns:
    c3, c4 = lines[n]["configs"]["cfg3"], lines[n]["configs"]["cfg4"]
    ag = "-" if c3.get("allgather_ms") is None else f"{c3['allgather_ms'] * 1e3:.0f} us / {c4['allgather_ms'] * 1e3:.0f} us"
    chk = "-" if not c4.get("class_shard_check") else f"max rel {c4['class_shard_check']['max_rel']:.1e}"
    out.append(f"| {n} | {c3['value']:,.0f} | {c3['ms_per_step']:.3f} | {c4['value']:,.0f} | {c4['ms_per_step']:.2f} | {ag} | {chk} |")
r1c4 = L("r01_bench_cfg4_n1.json")
out.append(f"\nRound 1, cfg4, N = 1: {r1c4['value']:,.0f} images/s.\n")
out += ["## Per-kernel time inside a step (CUDA events on the launching stream, kernels run one at a time; r02_bench_cfg2_n1.json `kernels`)\n",
        "| kernel | launches/step | us/launch | share of the serialised sum |", "|---|---|---|---|"]
for name, v in sorted(n1["kernels"].items(), key=lambda kv: -kv[1]["ms_total"]):
    out.append(f"| {name} | {v['launches'] / n1['steps']:.0f} | {v['ms_per_launch'] * 1e3:.1f} | {v['share'] * 100:.1f} % |")
rf = n1["roofline"]
out += [f"\n`gpu_launches` = {n1['gpu_launches']} over {n1['steps']} steps ({n1['gpu_launches'] // n1['steps']} per step, all from libschemahead.so; "
        "the timed loop replays them as one CUDA graph).\n",
        "## Rooflines (peaks: MEASURED_PEAKS.json -- HBM 6550 GB/s, bf16 / fp16 1388 TFLOP/s sustained)\n",
        "| stage | ms (events, whole stage incl. its small kernels) | bound | achieved | fraction of the peak |", "|---|---|---|---|---|"]
for name, v in n1["stages"].items():
    out.append(f"| {name} | {v['ms']:.3f} | {v['bound']} | {v['achieved']:.0f} {v['unit']} | {v['frac']:.2f} |")
out.append(f"\nDominant kernel (`roofline` of the line): {rf['kernel']} -- {rf['achieved']:.0f} TFLOP/s of fp32-accurate multiply-adds on the "
           f"visited tiles = {3 * rf['achieved']:.0f} TFLOP/s of fp16 MMAs (3 per product), frac {rf['frac']:.3f} of the sustained peak "
           f"(<= 1/3 by construction; tensor pipe {rf['tensor_pipe_frac']:.2f}); DRAM traffic per step of these launches "
           f"{rf['traffic'] / 1e6:.0f} MB ({rf['traffic_source']}).\n")
tr = L("r02_ncu_traffic.json")
out += ["## DRAM bytes per launch (`ncu --set full`, r02_ncu_traffic.json / r02_ncu_full_kernels.txt)\n", "| kernel | MB per launch (largest captured) | us under ncu |", "|---|---|---|"]
for k, v in tr.items():
    if isinstance(v, dict) and "dram_bytes_per_launch" in v:
        out.append(f"| {k} | {v['dram_bytes_per_launch'] / 1e6:.0f} | {v['us_per_launch_ncu']:.1f} |")
out += ["\n## Files\n",
        "* r02_bench_cfg2_n{1,2,4,8}.json -- `bench.py --gpus N --steps 20 --warmup 3` (N > 1 under torchrun), r02_bench_reference_cpu.json -- `--impl reference`",
        "* r02_launches_bench_cfg2.txt -- ncu launch list of `bench.py --steps 2 --warmup 1` (cold-cache, serialised: compare shares)",
        "* r02_ncu_full_kernels.txt, r02_ncu_traffic.json -- `ncu --set full` of the hot kernels (tools/ncu_summary.py)",
        "* r02_compute_sanitizer.txt -- memcheck (26 parity tests), racecheck and synccheck (smoke, wide fused path, fused class side): 0 errors / 0 hazards",
        "* r02_sass_mnemonics.txt -- tcgen05 / TMA (incl. `UTMALDG.2D.GATHER4`) / TMEM / cp.async mnemonic counts of the shipped library",
        "* r02_experiments.md -- variants measured this round, kept or not, with their numbers"]
open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote profiles/r02_summary.md")
