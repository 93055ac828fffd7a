"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Golden for BASELINE.json configs[2]'s MODEL at one sequence: the UNMODIFIED reference's
OPT-1.3B W6A6 block_fp forward (random init, seed 0, 1 x 2048 synthetic tokens, CPU fp32, bfp_6bit.toml)
-> tests/golden/opt13b_bfp6.npz.

Stored (small): loss, per-token log-partition, a 128 x 786 logits sample, and for EVERY decoder layer's output hidden state a
64 x 128 sampled slice, float64 checksums and the rms of the layer's update — the anchors of the noise-floor control and the
teacher-forced per-layer test in tests/test_gpu_parity_opt13b.py.  The 1.3 G weights are regenerated from the seed on the GPU box
(the mirror's seeded init equals the reference's bit for bit — asserted here; per-tensor checksums are stored and re-checked there).

Run in the authoring container only (needs /root/reference):  python oracle/gen_golden_opt13b.py      (~10 min on 8 cores)
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
OPT13B = dict(hidden_size=2048, num_hidden_layers=24, ffn_dim=8192, num_attention_heads=32, vocab_size=50272,
              max_position_embeddings=2048)
TOK_STRIDE, DIM_STRIDE = 32, 16           # hidden-state sample: rows ::32, columns ::16


def checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def main():
    models = ref_loader.load_models()
    qc = json.load(open(os.path.join(GOLD, "configs.json")))["raw"]["bfp_6bit.toml"]
    torch.manual_seed(0)
    ref = models.opt.OPTQuantizedForCausalLM(models.opt_cfg.OPTQuantizedConfig(quant_config=deepcopy(qc), **OPT13B)).eval()
    sums = checksums(ref.state_dict())                      # BEFORE the in-place PTQ overwrite of the first forward

    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM
    torch.manual_seed(0)
    ours = OPTQuantizedForCausalLM(OPTQuantizedConfig(quant_config=deepcopy(qc), tie_word_embeddings=False, **OPT13B)).eval()
    so, sr = ours.state_dict(), ref.state_dict()
    assert set(so) == set(sr) and all(torch.equal(so[k], sr[k]) for k in sr), "seeded init differs from the reference's"
    del ours, so, sr

    ids = torch.randint(0, OPT13B["vocab_size"], (1, SEQ), generator=torch.Generator().manual_seed(0))
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    with torch.no_grad():
        out = ref(input_ids=ids, labels=ids, output_hidden_states=True)
    dt = time.time() - t0
    logits = out.logits[0]
    hs = out.hidden_states                                  # input of layer 0 ... input of layer 23, then the final-LN output
    L = OPT13B["num_hidden_layers"]
    assert len(hs) == L + 1
    # hs[i] = input of layer i (i < L); the output of layer L-1 is only available before the final LayerNorm through a hook-free
    # identity: hs[L] is AFTER final_layer_norm, so layer outputs are stored for layers 0..L-2 (= hs[1..L-1]) plus the final-LN output.
    h_in = [h[0] for h in hs[:L]]
    arrs = {
        "input_ids": ids.numpy(),
        "loss": np.array(float(out.loss)),
        "logits_sub": logits[::16, ::64].numpy().copy(),
        "logits_row_lse": torch.logsumexp(logits.double(), -1).numpy(),
        "logits_std": np.array(float(logits.std())),
        "cpu_seconds": np.array(dt),
        "checksum_keys": np.array(list(sums.keys())),
        "checksum_vals": np.array(list(sums.values()), dtype=np.float64),
        # layer inputs 0..L-1 (input of layer i+1 == output of layer i) and the final-LayerNorm output
        "h_in_sub": np.stack([h[::TOK_STRIDE, ::DIM_STRIDE].numpy() for h in h_in]),
        "h_in_sum": np.array([[float(h.double().sum()), float(h.double().abs().sum())] for h in h_in]),
        "h_in_rms": np.array([float(h.double().pow(2).mean().sqrt()) for h in h_in]),
        "update_rms": np.array([float((h_in[i + 1] - h_in[i]).double().pow(2).mean().sqrt()) for i in range(L - 1)]),
        "h_final_sub": hs[L][0][::TOK_STRIDE, ::DIM_STRIDE].numpy().copy(),
        "sub_strides": np.array([TOK_STRIDE, DIM_STRIDE]),
    }
    np.savez_compressed(os.path.join(GOLD, "opt13b_bfp6.npz"), **arrs)
    print(f"opt13b_bfp6: loss {float(out.loss):.6f}  ({dt:.1f} s on {os.cpu_count()} cores = {SEQ / dt:.1f} tokens/s)")


if __name__ == "__main__":
    main()
