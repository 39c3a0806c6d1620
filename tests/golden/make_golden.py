"""Generates tests/golden/path_v1.pt: frozen outputs of the fp32 CPU oracle for fixed seeds.

The reference ships no fixtures (/root/reference/README.md:11-16), so these vectors freeze the
oracle *after* it passed the analytic KATs (tests/test_oracle_kat.py); they guard the oracle against
silent drift and give the GPU parity tests a target that does not need the oracle at run time.
Inputs are regenerated from seeds by spec.synthetic_inputs; only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import styletts_zs_b200 as stz  # noqa: E402
from oracle.model import OraclePath  # noqa: E402

CASES = {
    # name: (B, T, steps, sampler, var_len, seed, cfg_scale)
    "student1": (2, 32, 1, 0, None, 1234, 2.0),
    "student4": (2, 32, 4, 0, None, 1235, 2.0),
    "teacher2": (2, 32, 2, 1, None, 1236, 2.0),
    "varlen_student2": (3, 40, 2, 0, (8, 40), 1237, 1.5),
}


def run_case(o, cfg, case):
    B, T, steps, sampler, var_len, seed, scale = case
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=sampler, seed=seed, var_len=var_len)
    z = o.sample_style(inp["text_emb"], inp["prompt_feats"], steps, scale, text_mask=inp["text_mask"],
                       noise=inp["noise"], sampler=sampler)
    dur, pre = o.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"], return_presum=True)
    return dict(style=z, dur=dur, presum=pre)


if __name__ == "__main__":
    torch.set_num_threads(8)
    cfg = stz.DEFAULT
    o = OraclePath(cfg, stz.init_weights(cfg, 0))
    out = {name: run_case(o, cfg, case) for name, case in CASES.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "path_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
