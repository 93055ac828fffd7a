"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Golden for BASELINE.json configs[0]: the UNMODIFIED reference's OPT-125M W6A6 block_fp
forward (random init, seed 0, 1 x 2048 synthetic tokens, CPU fp32) -> tests/golden/opt125m_bfp6.npz.

The 125 M weights are not stored: every parameter of the reference model built under torch.manual_seed(0) equals, bit for bit,
the parameter of llm_mixed_q_b200's OPTQuantizedForCausalLM built the same way with tie_word_embeddings=False (checked here,
and re-checked on the GPU box through the per-tensor checksums saved below), so the test regenerates them from the seed.

Run in the authoring container only (needs /root/reference):  python oracle/gen_golden_opt125m.py
"""
import json
import os
import sys
import time
from copy import deepcopy

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEQ = 2048


def checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def main():
    models = ref_loader.load_models()
    qc = json.load(open(os.path.join(GOLD, "configs.json")))["raw"]["bfp_6bit.toml"]
    torch.manual_seed(0)
    ref = models.opt.OPTQuantizedForCausalLM(models.opt_cfg.OPTQuantizedConfig(quant_config=deepcopy(qc))).eval()
    sums = checksums(ref.state_dict())                      # BEFORE the in-place PTQ overwrite of the first forward

    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM
    torch.manual_seed(0)
    ours = OPTQuantizedForCausalLM(OPTQuantizedConfig(quant_config=deepcopy(qc), tie_word_embeddings=False)).eval()
    so, sr = ours.state_dict(), ref.state_dict()
    assert set(so) == set(sr) and all(torch.equal(so[k], sr[k]) for k in sr), "seeded init differs from the reference's"
    del ours

    ids = torch.randint(0, 50272, (1, SEQ), generator=torch.Generator().manual_seed(0))
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    with torch.no_grad():
        out = ref(input_ids=ids, labels=ids)
    dt = time.time() - t0
    logits = out.logits[0]
    arrs = {
        "input_ids": ids.numpy(),
        "loss": np.array(float(out.loss)),
        "logits_sub": logits[::16, ::64].numpy().copy(),                       # 128 x 786 sample of the 2048 x 50272 logits
        "logits_row_lse": torch.logsumexp(logits.double(), -1).numpy(),        # per-token log-partition (what the loss sums)
        "logits_std": np.array(float(logits.std())),
        "cpu_seconds": np.array(dt),
        "checksum_keys": np.array(list(sums.keys())),
        "checksum_vals": np.array(list(sums.values()), dtype=np.float64),
    }
    np.savez_compressed(os.path.join(GOLD, "opt125m_bfp6.npz"), **arrs)
    print(f"opt125m_bfp6: loss {float(out.loss):.6f}  ({dt:.1f} s on {os.cpu_count()} cores = {SEQ / dt:.1f} tokens/s)")


if __name__ == "__main__":
    main()
