"""Golden on-disk outputs of the UNMODIFIED reference: event log, plateau exit, data_collection/data.txt
column and createLogFile text for the golden cases below, written to tests/golden/outputs.json.  Run here (needs
/root/reference):  python tests/golden/make_outputs_golden.py"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import ref_harness as rh  # noqa: E402

CASES = {"c1_default_moveeat_mt": 330, "c1_default_moveeat_keyed": 150, "c2_pollution_mt": 150}
out = {}
for name, limit in CASES.items():
    g = common.load_golden(name)
    ref = rh.ReferenceRun(g["settings"], rng=g["rng"])
    steps = 0
    while steps < limit and ref.exited_at is None:
        ref.step()
        steps += 1
    with rh._cwd(ref.scratch):
        ref.view.createLogFile()
        text = open(glob.glob(os.path.join(ref.scratch, "logs", "log_*.txt"))[0]).read()
    dp = os.path.join(ref.scratch, "data_collection", "data.txt")
    out[name] = dict(steps=steps, exited_at=ref.exited_at, log=list(ref.view.log), log_file_text=text,
                     data_txt=open(dp).read() if os.path.exists(dp) else None, stats=ref.view.stats,
                     pop_at_100=ref.view.pop_at_100, gini_at_100=ref.view.gini_at_100)
    print(name, "steps", steps, "exited_at", ref.exited_at, "log lines", len(ref.view.log), "data", out[name]["data_txt"])
with open(os.path.join(HERE, "outputs.json"), "w") as f:
    json.dump(out, f)
print("wrote", os.path.getsize(os.path.join(HERE, "outputs.json")), "bytes")
