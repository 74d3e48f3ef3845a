"""Runs bench.py with the given extra arguments and prints ms/step plus the per-launch time of each contraction class.
    python scripts/kernel_ms.py --config c4 --T 524288 --steps 3 --warmup 3 --no-e2e --no-cpu
"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *sys.argv[1:]], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(1)
d = json.loads(line[-1])
per = {k: round(v["total_ms"] / max(v["launches"], 1), 2) for k, v in d["roofline"]["kernel_ms"].items()}
print(os.environ.get("CMF_TAG", ""), "ms/step", round(d["ms_per_step"], 2), "direct-loss it/s", d.get("value_direct_loss"),
      per, "clk", d["clocks"]["sm_mhz"], "loss", d["loss"]["final"])
