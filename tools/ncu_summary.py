"""Summarise .ncu-rep files (ncu --set full) into the few numbers profiles/ quotes: duration, DRAM bytes, DRAM %, tensor-pipe %,
SM busy %, achieved occupancy, registers, top stall reasons."""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "inst"), ("sm__cycles_elapsed.max", "cycles"),
    ("l1tex__t_bytes.sum", "l1_bytes"), ("lts__t_bytes.sum", "l2_bytes"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [(i, h) for i, h in enumerate(hdr) if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    tens = [(i, h) for i, h in enumerate(hdr) if "tensor" in h and ("pct" in h)]
    for r in rows[2:]:
        d = {}
        for name, short in WANT:
            if name in hdr:
                i = hdr.index(name)
                d[short] = f"{r[i]} {units[i]}".strip()
        print(d)
        tl = sorted(((float(r[i].replace(",", "") or 0), h) for i, h in tens if r[i] not in ("", "no data", "n/a")), reverse=True)[:3]
        print("   tensor:", [(round(v, 1), h.split(".")[0][-60:]) for v, h in tl])
        sl = sorted(((float(r[i].replace(",", "") or 0), h) for i, h in stall if r[i] not in ("", "no data", "n/a")), reverse=True)[:4]
        print("   stalls:", [(round(v, 1), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")) for v, h in sl])


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        main(p)
