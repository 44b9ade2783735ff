"""Average DRAM traffic per GEMM launch of one training step, from
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm16 --csv \
        --log-file gpurun_out/gemm_traffic.csv python bench.py --ncu
Writes profiles/gemm_traffic.json, which bench.py reports as roofline.traffic (bytes per launch, like `achieved`).
Usage: python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv <tag>"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, tag):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    iid, im, iu, iv = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = {}
    for r in data:
        if len(r) <= iv or not r[im].startswith("dram__bytes"):
            continue
        per.setdefault(r[iid], 0.0)
        per[r[iid]] += float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
    n = len(per)
    tot = sum(per.values())
    out = {"tag": tag, "launches": n, "dram_bytes_per_step": tot, "traffic_bytes_per_launch": tot / max(n, 1),
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every gemm16* launch of one step "
                     "(bench.py --ncu, batch 32)"}
    json.dump(out, open("profiles/gemm_traffic.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r01")
