"""Summarise the round-2 ncu captures (gpurun_out/r2_*.ncu-rep, r2_launches.csv) into profiles/:
   r2_launches.csv (copy), r2_ncu_summary.md (per-kernel metrics + share of the step), r2_ncu_gae.json (dram bytes/launch)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

KEYS = [
    ("gpu__time_duration.sum", "duration (us)"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block (KB)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (smem)"), ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active (%)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (%)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe active (%)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput (% of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
]
STALLS = ["long_scoreboard", "no_instruction", "wait", "short_scoreboard", "barrier", "not_selected", "mio_throttle", "math_pipe_throttle",
          "branch_resolving", "dispatch_stall", "lg_throttle"]


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def section(name, rep, note=""):
    d = raw(rep)
    if d is None:
        return f"### {name}\n(capture missing)\n\n", None
    out = [f"### {name}", f"kernel: `{d.get('Kernel Name', ('?', ''))[0]}`  " + note, "", "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in d:
            out.append(f"| {label} | {d[k][0]} {d[k][1]} |")
    st = []
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in d:
            st.append(f"{s} {float(d[k][0]):.2f}")
    out += ["", "stall cycles per issued instruction: " + ", ".join(st), ""]
    return "\n".join(out) + "\n", d


def launches():
    p = os.path.join(G, "r2_launches.csv")
    if not os.path.exists(p):
        return "(launch list missing)\n"
    rows = [r for r in csv.reader(open(p)) if r and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        try:
            v = float(r[-1].replace(",", ""))
        except ValueError:
            continue
        unit = r[-2]
        us = v / 1e3 if unit.startswith("ns") or unit == "nsecond" else (v if unit.startswith("us") else v * 1e3 if unit.startswith("ms") else v)
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total (us) | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f"| `{k[:90]}` | {n} | {us:.1f} | {100 * us / max(tot, 1e-9):.1f} % |")
    with open(os.path.join(OUT, "r2_launches.csv"), "w") as f:
        f.write(open(p).read())
    return "\n".join(out) + f"\n\nsum over the captured launches: {tot / 1e3:.2f} ms (cold-cache, serialised by ncu: shares, not absolutes, are comparable with bench.py)\n"


def main():
    os.makedirs(OUT, exist_ok=True)
    md = ["# Round-2 ncu summaries (B200, `--clock-control none`)", "",
          "Captured by `tools/profile_r2.sh` under gpurun; raw reports stay in gpurun_out/ (scratch).", "",
          "## Launch list of `bench.py --steps 2 --warmup 3` (eager launches, `--metrics gpu__time_duration.sum`)", "", launches(), "",
          "## Full captures (`--set full`, one launch each, C2: CartPole-v1, 4096 envs, T = 128)", ""]
    s, d = section("ppo_fwdbwd_tc_kernel — the update (dominant kernel)", os.path.join(G, "r2_tc_update.ncu-rep"),
                   "524 288 rows, whole-buffer minibatch (TMA staging), policy + critic nets, 2 CTAs/SM")
    md.append(s)
    s, _ = section("rollout_cartpole_q5_kernel — fused rollout, 128 steps", os.path.join(G, "r2_tc_rollout.ncu-rep"),
                   "32 CTAs x 128 envs x (4 forward + 2 env) threads; a chain of 128 dependent steps, latency / issue bound by design")
    md.append(s)
    s, _ = section("critic_values_tc_kernel — value pass over (T+1)*B rows", os.path.join(G, "r2_tc_critic.ncu-rep"))
    md.append(s)
    s, g = section("gae_scan_kernel at the >= 1 GB shape (T = 128, B = 2^21, 24 B/element)", os.path.join(G, "r2_gae1g.ncu-rep"))
    md.append(s)
    if g is not None:
        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd, wr = to_bytes(*g["dram__bytes_read.sum"]), to_bytes(*g["dram__bytes_write.sum"])
        json.dump({"kernel": g.get("Kernel Name", ("?", ""))[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "algorithmic_bytes_per_launch": 128 * (1 << 21) * 24, "source": "ncu --set full, tools/gae_1gb.py (tools/profile_r2.sh)"},
                  open(os.path.join(OUT, "r2_ncu_gae.json"), "w"), indent=1)
        md.append(f"dram read + write = {(rd + wr) / 1e9:.3f} GB vs {128 * (1 << 21) * 24 / 1e9:.3f} GB algorithmic -> ratio {(rd + wr) / (128 * (1 << 21) * 24):.3f}\n")
    open(os.path.join(OUT, "r2_ncu_summary.md"), "w").write("\n".join(md))
    print("wrote profiles/r2_ncu_summary.md")


if __name__ == "__main__":
    main()
