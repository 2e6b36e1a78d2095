#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the reference's own known-answer fixtures.

The reference pins its classic-control envs with Gymnasium-generated JSON
(/root/reference/python_tests/{cartpole,mountain_car}/{inputs,output}.json, replayed
by src/testing.rs:34-146).  /root/reference does not exist on the GPU box, so the
vectors are committed here, one compact file per env, with provenance.

Run in the build container only:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os

REF = "/root/reference/python_tests"
HERE = os.path.dirname(os.path.abspath(__file__))

for folder in ("cartpole", "mountain_car"):
    src_in = os.path.join(REF, folder, "inputs.json")
    src_out = os.path.join(REF, folder, "output.json")
    raw_in = open(src_in, "rb").read()
    raw_out = open(src_out, "rb").read()
    actions = json.loads(raw_in)
    outputs = json.loads(raw_out)
    assert len(actions) == len(outputs) == 100
    doc = {
        "source": {
            "inputs": f"python_tests/{folder}/inputs.json",
            "output": f"python_tests/{folder}/output.json",
            "inputs_sha256": hashlib.sha256(raw_in).hexdigest(),
            "output_sha256": hashlib.sha256(raw_out).hexdigest(),
            "generator": "python_tests/test_gym.py (Gymnasium, zero-state resets, 100 random actions)",
            "replayed_by": "src/testing.rs:34-146 (teacher forced, tolerance 1e-4)",
        },
        "actions": actions,
        "observation": [o["observation"] for o in outputs],
        "reward": [o["reward"] for o in outputs],
        "done": [o["done"] for o in outputs],
        "truncated": [o["truncated"] for o in outputs],
    }
    dst = os.path.join(HERE, f"{folder}_gymnasium.json")
    with open(dst, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote", dst, os.path.getsize(dst), "bytes")
    # the same vectors as flat text for the C++ test: action obs... reward done truncated (floats as %.9g)
    txt = os.path.join(HERE, f"{folder}_gymnasium.txt")
    with open(txt, "w") as f:
        f.write(f"# from python_tests/{folder}/{{inputs,output}}.json; columns: action obs[{len(outputs[0]['observation'])}] reward done truncated\n")
        for a, o in zip(actions, outputs):
            cols = [str(a)] + ["%.9g" % x for x in o["observation"]] + ["%.9g" % o["reward"], str(int(o["done"])), str(int(o["truncated"]))]
            f.write(" ".join(cols) + "\n")
    print("wrote", txt)
