"""Generates tests/golden/*.json from the oracle (the reference cannot run here: no cargo, no weights).
The fixtures freeze the oracle's outputs on the reference's own weight-free sampler fixture
(benches/sampling.rs:12-64) so that later oracle edits cannot silently change them.
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import sampling as s  # noqa: E402

F = np.float32
i = np.arange(3072, dtype=F)
logits = (np.sin(i * F(0.1)) * F(5.0)).astype(F)[None]
cases = []
for top_k, top_p in [(50, 1.0), (0, 0.5), (0, 0.9), (0, 0.95), (50, 0.9)]:
    cfg = s.GenerationConfig(temperature=0.9, top_k=top_k, top_p=top_p, repetition_penalty=1.0)
    ctx = s.SamplingContext(42)
    cases.append(dict(top_k=top_k, top_p=top_p, tokens=[int(s.sample(logits, cfg, ctx)[0]) for _ in range(32)]))
ctx = s.SamplingContext(42)
json.dump(dict(cases=cases, pcg_seed42_u32=[ctx.next_u32() for _ in range(8)]),
          open(os.path.join(HERE, "sampler_fixture.json"), "w"), indent=1)
print("wrote sampler_fixture.json")
