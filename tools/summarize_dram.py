"""ncu csv with dram__bytes_read.sum / dram__bytes_write.sum per launch (one bench step, GEMM kernels only) ->
profiles/gemm_traffic.json  {dram_bytes_per_launch, launches, total_bytes, source} (read by bench.py: roofline.traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
        -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram.csv python bench.py --profile-step --no-cpu-baseline
    python tools/summarize_dram.py gpurun_out/gemm_dram.csv profiles/gemm_traffic.json "r1d ncu pass" 64 bench
"""
import csv
import json
import sys


def main(path, out, source, chunks, shape):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = {}
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"].lower()
        v *= {"byte": 1, "bytes": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        per[row["ID"]] = per.get(row["ID"], 0.0) + v
    n = len(per)
    tot = sum(per.values())
    res = {"dram_bytes_per_launch": tot / max(n, 1), "launches": n, "total_bytes": tot, "source": source,
           "chunks": int(chunks), "shape": shape}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    main(*sys.argv[1:6])
