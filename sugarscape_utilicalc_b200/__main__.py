"""Headless run of a settings.json on the engine -- what `python sugarscape.py` does without its Tk window:

    python -m sugarscape_utilicalc_b200 [settings.json] [--steps N] [--rng mt|keyed] [--no-events] [--quiet]

`random.seed(seed)`, the world of `__main__` (sugarscape.py:1485-1555) built on the device, then `updateGame` until
the plateau rule exits (sugarscape.py:679-682, data_collection/data.txt gets its column) or N steps have run; with
`view.create_log_on_exit` the log file of `createLogFile` is written to ./logs.  MT-exact mode reproduces the
reference's run of the same settings.json bit for bit.
"""
import argparse
import os
import sys
import time

from . import config_from_settings, init_params_from_settings, load_settings, outputs, rule_check
from .engine import Engine


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m sugarscape_utilicalc_b200")
    ap.add_argument("settings", nargs="?", default="settings.json")
    ap.add_argument("--steps", type=int, default=None, help="stop after N timesteps (default: run until the plateau rule)")
    ap.add_argument("--rng", choices=["mt", "keyed"], default="mt")
    ap.add_argument("--no-events", action="store_true", help="skip the per-step death log (one id read-back per step)")
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args(argv)
    settings = load_settings(args.settings)
    rule_check(settings)
    seed = settings["env"]["random_seed"]
    if settings["env"].get("is_random"):
        raise SystemExit("env.is_random draws the seed from the host's entropy; set env.random_seed instead")
    if not args.quiet:
        print("Seed: ", seed)                                        # sugarscape.py:1489
    eng = Engine(config_from_settings(settings, rng_mode=args.rng))
    eng.seed_mt19937(seed)
    eng.init_world(init_params_from_settings(settings))
    os.makedirs("data_collection", exist_ok=True)
    rec = outputs.RunRecorder(settings, eng, rng=args.rng, events=not args.no_events)
    t0 = time.perf_counter()
    code = 0
    try:
        while args.steps is None or rec.iteration < args.steps:
            rec.step()
            if not args.quiet and settings["view"].get("update_rate", 1) and rec.iteration % 50 == 0:
                print("Iteration = %d | Population = %d" % (rec.iteration, rec.population[-1]))
    except outputs.PlateauReached:
        if not args.quiet:
            print("population unchanged for 200 steps at iteration %d: data_collection/data.txt updated" % rec.iteration)
    dt = time.perf_counter() - t0
    if settings["view"].get("create_log_on_exit"):
        path = rec.create_log_file()
        if not args.quiet:
            print("log file:", path)
    if not args.quiet:
        print("%d steps, population %d, gini %.4f, %.1f steps/s" % (rec.iteration, rec.population[-1] if rec.population else 0,
                                                                    rec.gini[-1] if rec.gini else 0.0, rec.iteration / max(dt, 1e-9)))
    eng.close()
    return code


if __name__ == "__main__":
    sys.exit(main())
