"""ncu report(s) (--set full) -> one markdown table of the metrics SURVEY.md §8(d) asks for, per captured launch:
duration, tensor-pipe utilisation, achieved DRAM bandwidth (+ read / write bytes), SM throughput, issue-slot use,
registers, achieved occupancy.

    python tools/ncu_key_metrics.py gpurun_out/prof_layer_r1.ncu-rep [more.ncu-rep ...] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1),
    ("dram__bytes.sum.per_second", "DRAM GB/s", 1e-9),
    ("dram__bytes_read.sum", "DRAM rd MB", 1e-6),
    ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte/second": 1, "Kbyte/second": 1e3,
              "Mbyte/second": 1e6, "Gbyte/second": 1e9, "Tbyte/second": 1e12, "byte/s": 1, "Kbyte/s": 1e3,
              "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "nsecond": 1, "usecond": 1e3,
              "msecond": 1e6, "second": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if not l.startswith("==")]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    hdr = next(rd)
    units = next(rd)
    for r in rd:
        yield {h: (v, u) for h, v, u in zip(hdr, r, units)}


def main(paths):
    print("| kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for p in paths:
        for r in rows_of(p):
            name = r["Kernel Name"][0].replace("void ", "").replace("rvb::", "").split("(")[0][:44]
            cells = []
            for key, _, scale in COLS:
                v, u = r.get(key, ("", ""))
                try:
                    x = float(v.replace(",", "")) * UNIT_SCALE.get(u, 1) * scale
                    cells.append(f"{x:.1f}" if abs(x) < 1e5 else f"{x:.3g}")
                except ValueError:
                    cells.append("-")
            print(f"| `{name}` | {r['Grid Size'][0]} x {r['Block Size'][0]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
