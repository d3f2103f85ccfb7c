"""A/B helper: runs bench.py for combinations of environment knobs and prints ms / step.
Usage: python profiles/ab_env.py WORKLOAD STEPS REPS KNOB=v1,v2 [--streams=1,3]"""
import itertools
import json
import os
import subprocess
import sys

wl, steps, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
knobs, streams = [], ["0"]
for a in sys.argv[4:]:
    if a.startswith("--streams="):
        streams = a.split("=", 1)[1].split(",")
    else:
        k, v = a.split("=", 1)
        knobs.append((k, v.split(",")))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(reps):
    for combo in itertools.product(*[v for _, v in knobs], streams):
        env = dict(os.environ)
        for (k, _), v in zip(knobs, combo[:-1]):
            if v != "-":
                env[k] = v
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", wl, "--streams", combo[-1], "--steps", steps, "--warmup", "3",
                              "--no-cpu-baseline", "--no-extra"], env=env, capture_output=True, text=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        print(" ".join(f"{k}={v}" for (k, _), v in zip(knobs, combo[:-1])), f"streams={combo[-1]}", f"{d['ms_per_step']:.3f} ms/step", flush=True)
