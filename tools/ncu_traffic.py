"""profiles/traffic.json from `ncu --set full ... --page raw --csv` exports of ONE timestep each:
    python tools/ncu_traffic.py c5=profiles/r02_ncu_full_c5_step.csv c4=... c5u=...
Per workload and kernel class (move = conflict graph + dependency rounds / move passes + settle, grow, diffuse) the sum of
dram__bytes_read.sum + dram__bytes_write.sum over the launches of the captured step; `kernels_sha` ties the figures to
the CUDA sources they were captured on (bench.py reports `traffic: null` for any other sources)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
CLASSES = (("move", ("ss_dr_build", "ss_dr_rounds", "ss_dr_seg0", "ss_dr_begin", "ss_dr_check", "ss_move_", "ss_settle", "ss_prepare")),
           ("grow", ("ss_dr_grow", "ss_grow", "ss_seasons")), ("diffuse", ("ss_diffuse",)))


def main():
    out = {"kernels_sha": bench.kernels_sha(), "entries": {}}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        it = h.index("gpu__time_duration.sum")
        ent = {}
        for r in rows[2:]:
            name = r[ik]
            for cls, prefixes in CLASSES:
                if any(p in name for p in prefixes):
                    e = ent.setdefault(cls, {"dram_bytes_per_step": 0.0, "launches": 0, "ncu_ms": 0.0, "kernels": []})
                    e["dram_bytes_per_step"] += float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
                    e["launches"] += 1
                    e["ncu_ms"] += float(r[it]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[it], 1e-6)
                    short = name.split("(")[0].replace("void ", "")
                    if short not in e["kernels"]:
                        e["kernels"].append(short)
        for cls, e in ent.items():
            e["source"] = "%s (ncu --set full, one timestep, %d launches)" % (os.path.relpath(path, ROOT), e["launches"])
        ent["agent"] = ent.get("move", None)        # the secondary workloads of bench.py call the class "agent phase"
        if ent["agent"] is None:
            del ent["agent"]
        out["entries"][key] = ent
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
