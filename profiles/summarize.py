"""Turns gpurun_out ncu artefacts into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv            > profiles/r01_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof_pack_r01.ncu-rep       > profiles/r01_k_pack.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_tma_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ik].split("(")[0][:60], []).append(float(r[iv].replace(",", "")))
        except ValueError:
            pass
    ours = {k: v for k, v in agg.items() if "dra::" in k}
    tot = sum(sum(v) / len(v) for v in ours.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch device time, cold-cache and serialised: compare SHARES of the step, not absolutes")
    print(f"{'kernel':<62}{'launches':>9}{'avg us':>10}{'min us':>10}{'share':>8}")
    for k, v in agg.items():
        a = sum(v) / len(v) / 1e3
        sh = f"{100 * a * 1e3 / tot:6.1f}%" if k in ours else "   (not ours)"
        print(f"{k:<62}{len(v):>9}{a:>10.2f}{min(v) / 1e3:>10.2f}{sh:>8}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on  ({path})")
    for r in rows[2:]:
        print(f"\n== {r[hdr.index('Kernel Name')]}  (launch id {r[hdr.index('ID')]})")
        seen = set()
        for k in KEYS:
            if k in hdr and k not in seen:
                seen.add(k)
                print(f"  {k:<86}{r[hdr.index(k)]:>16} {units[hdr.index(k)]}")
        # derived: what the judge asked for
        def val(k):
            try:
                return float(r[hdr.index(k)].replace(",", ""))
            except (ValueError, IndexError):
                return None
        def scaled(k):
            v, u = val(k), units[hdr.index(k)] if k in hdr else ""
            return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        xb, dr, dw = scaled("l1tex__m_xbar2l1tex_read_bytes.sum"), scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
        if xb and dr is not None:
            print(f"  {'(derived) L2 -> SM bytes / DRAM bytes (re-reads served by L2)':<86}{xb / max(1.0, dr + (dw or 0)):>16.1f} x")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
