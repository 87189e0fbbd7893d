"""Extracts the ASCII golden images of the reference's 2D render tests
(fidget/tests/pixel_render.rs:70-364) into pixel_render.json.

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/extract_pixel_goldens.py /root/reference
The JSON is committed; the tests only read the JSON.
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "fidget/tests/pixel_render.rs")).read()
out = {}
for fn in re.finditer(r"fn (check_\w+)<", src):
    start = fn.end()
    nxt = src.find("\nfn ", start)
    body = src[start: nxt if nxt > 0 else len(src)]
    for m in re.finditer(r'const (EXPECTED\w*): &str = "\n(.*?)";', body, re.S):
        rows = [r.strip() for r in m.group(2).split("\n") if r.strip()]
        line = src[:start + m.start()].count("\n") + 1
        out[f"{fn.group(1)}:{m.group(1)}"] = {"rows": rows, "source": f"fidget/tests/pixel_render.rs:{line}"}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pixel_render.json")
json.dump(out, open(path, "w"), indent=1)
print({k: (len(v["rows"]), len(v["rows"][0])) for k, v in out.items()})
