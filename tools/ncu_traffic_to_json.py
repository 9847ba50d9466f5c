"""gpurun_out/ncu_spmm_w{1,4}_r02.csv (written by tools/ncu_spmm_traffic.sh) -> profiles/spmm_traffic.json + a summary table."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].strip('"').isdigit()]
    out = {}
    for r in rows:
        out.setdefault(r[0], {"kernel": r[4]})[r[-3]] = (r[-1], r[-2])
    return list(out.values())


def num(v):
    return float(v[0].replace(",", ""))


def main():
    res, lines = {}, ["# ncu metrics of the SpMM launches behind bench.py's roofline (round 2; tools/ncu_spmm_traffic.sh)", ""]
    for world, name in ((1, "ncu_spmm_w1_r02.csv"), (4, "ncu_spmm_w4_r02.csv")):
        p = os.path.join(ROOT, "gpurun_out", name)
        if not os.path.exists(p):
            continue
        launches = read(p)
        dram = sum(num(k["dram__bytes_read.sum"]) + num(k["dram__bytes_write.sum"]) for k in launches)
        unit = launches[0]["dram__bytes_read.sum"][1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        res[str(world)] = {"dram_bytes_per_launch": dram * scale, "kernel_launches_summed": len(launches),
                           "source": f"ncu --metrics dram__bytes_*.sum --clock-control none, profiles/{name} "
                                     "(tools/ncu_spmm_traffic.sh, round 2)"}
        lines.append(f"## world {world}: {len(launches)} kernel launch(es) per logical SpMM")
        for k in launches:
            lines.append("")
            lines.append(f"`{k['kernel'][:90]}`")
            for m, v in k.items():
                if m != "kernel":
                    lines.append(f"* {m}: {v[0]} {v[1]}")
        lines.append("")
    if res:
        res["dram_bytes_per_launch"] = res.get("1", {}).get("dram_bytes_per_launch")      # round-1 key (world 1)
        res["source"] = res.get("1", {}).get("source")
        with open(os.path.join(ROOT, "profiles", "spmm_traffic.json"), "w") as f:
            json.dump(res, f, indent=1)
        with open(os.path.join(ROOT, "gpurun_out", "ncu_spmm_r02_summary.md"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
