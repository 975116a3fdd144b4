"""Record the defaults of the reference driver's command line (nerf_mae/run_swin_mae3d.py:41-313 parse_args) into
tests/golden/ref_cli_defaults.json.  Test infrastructure; build container only.  The reference driver itself does not import here
(torchmetrics / matplotlib are missing), so only its parse_args function is executed, from its own source text."""
import json
import os
import sys

SRC = "/root/reference/nerf_mae/run_swin_mae3d.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_cli_defaults.json")

if __name__ == "__main__":
    src = open(SRC).read()
    start = src.index("def parse_args")
    end = src.index("\n\n\n", start)
    ns = {}
    exec("import argparse\n" + src[start:end], ns)
    sys.argv = ["run_swin_mae3d.py"]
    defaults = vars(ns["parse_args"]())
    with open(OUT, "w") as f:
        json.dump(defaults, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(defaults), "flags")
