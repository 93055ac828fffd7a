"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Golden-vector generator.  Runs the UNMODIFIED reference (imported read-only from
/root/reference/src through oracle/ref_loader.py) in the authoring container and
writes small fixtures to tests/golden/:

  quantizers.npz + manifest.json   inputs/outputs (int32 bit patterns) of every quantizer on
                                   the hot path over the edge-case matrix of SURVEY.md §4.1
  hashed.json                      sha256 of reference outputs on larger seeded inputs
  consumers.npz                    LinearBlock* / matmul_* / bmm_* reference outputs
  configs.json                     raw + reference-parsed quant configs (every shipped TOML)

Usage (authoring container only):  python oracle/gen_golden.py
The GPU box never runs this; tests read the committed fixtures.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_CFG_DIR = "/root/reference/experiments/emnlp/configs/quantization"


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.int32).numpy().copy()


# ---------------------------------------------------------------------------
# input generators (deterministic; stored in the fixture anyway)
# ---------------------------------------------------------------------------


def rs(seed):
    return np.random.RandomState(seed)


def gen_normal(shape, sigma, seed):
    return torch.from_numpy((rs(seed).standard_normal(shape) * sigma).astype(np.float32))


def gen_softmax(shape, seed):
    """post-softmax rows with a causal mask → exact zeros and all-zero blocks."""
    s = torch.from_numpy(rs(seed).standard_normal(shape).astype(np.float32)) * 3
    n = shape[-1]
    mask = torch.triu(torch.ones(shape[-2], n, dtype=torch.bool), diagonal=1)
    s = s.masked_fill(mask, torch.finfo(torch.float32).min)
    return torch.softmax(s, dim=-1)


def gen_cliffs(shape, seed):
    """values 2^k (1 + d 2^-23), d in [-4, 8]: the ceil/floor/round(log2) cliffs."""
    r = rs(seed)
    n = int(np.prod(shape))
    k = r.randint(-30, 31, size=n)
    d = r.randint(-4, 9, size=n)
    half = r.randint(0, 3, size=n) == 0          # some near sqrt(2)*2^k for the block_log round cliff
    base = np.where(half, np.float32(np.sqrt(2.0)), np.float32(1.0)).astype(np.float32)
    v = (base.view(np.int32) + d.astype(np.int32)).view(np.float32) * np.exp2(k).astype(np.float32)
    sgn = np.where(r.randint(0, 2, size=n) == 0, -1.0, 1.0).astype(np.float32)
    return torch.from_numpy((v * sgn).astype(np.float32).reshape(shape))


def gen_tiny(shape, seed):
    """|x| around the 1e-8 passthrough threshold and the 1e-9 epsilon, ±0, exact ±1e-9f."""
    r = rs(seed)
    n = int(np.prod(shape))
    pool = np.array(
        [0.0, -0.0, 1e-9, -1e-9, 1e-8, -1e-8, 9.9e-9, 1.01e-8, -1.01e-8, 5e-9, -5e-9, 2e-8, 1e-7, -3e-8, 1e-10, 1e-12,
         1e-38, -1e-38, 1e-41, 3e-5, -2e-3, 0.5, -1.0],
        dtype=np.float32,
    )
    return torch.from_numpy(pool[r.randint(0, len(pool), size=n)].reshape(shape))


def gen_zero_blocks(shape, sigma, seed):
    x = gen_normal(shape, sigma, seed)
    flat = x.reshape(-1, shape[-1])
    flat[::3] = 0                                   # whole rows zero → all-zero blocks
    flat[:, : min(16, shape[-1])] = 0               # first block of every row zero
    return flat.reshape(shape).clone()


def gen_extreme(shape, seed):
    r = rs(seed)
    n = int(np.prod(shape))
    k = r.randint(-140, 128, size=n)
    v = np.ldexp(r.uniform(1, 2, size=n), k).astype(np.float32)
    sgn = np.where(r.randint(0, 2, size=n) == 0, -1.0, 1.0).astype(np.float32)
    return torch.from_numpy((v * sgn).reshape(shape))


FORMATS = [
    ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127)),
    ("block_fp", dict(width=4, exponent_width=8, exponent_bias=None)),
    ("block_fp", dict(width=8, exponent_width=8, exponent_bias=127)),
    ("block_fp", dict(width=2, exponent_width=8, exponent_bias="NA->None")),
    ("block_fp", dict(width=5, exponent_width=4, exponent_bias=3)),
    ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8)),
    ("block_minifloat", dict(width=4, exponent_width=2, exponent_bias_width=8)),
    ("block_minifloat", dict(width=6, exponent_width=3, exponent_bias_width=2)),
    ("block_log", dict(width=8, exponent_bias_width=8)),
    ("block_log", dict(width=4, exponent_bias_width=8)),
    ("block_log", dict(width=5, exponent_bias_width=3)),
    ("minifloat_denorm", dict(width=8, exponent_width=4, exponent_bias=None)),
    ("minifloat_denorm", dict(width=4, exponent_width=2, exponent_bias=None)),
    ("minifloat_denorm", dict(width=6, exponent_width=3, exponent_bias=5)),
    ("minifloat_ieee", dict(width=8, exponent_width=4, exponent_bias=None)),
    ("integer", dict(width=8, frac_width=7)),
]
BLOCKED = ("block_fp", "block_minifloat", "block_log")

# (tag, tensor builder, skip_first_dim)
def layouts():
    yield "bias1d_48", (48,), False
    yield "bias1d_10", (10,), False            # shorter than a block
    yield "act2d_7x40", (7, 40), True          # ragged last block
    yield "w2d_9x40", (9, 40), False
    yield "w2d_32x64", (32, 64), False
    yield "act3d_3x5x40", (3, 5, 40), True
    yield "act3d_2x16x64", (2, 16, 64), True


def inputs_for(shape, base_seed):
    yield "n1e-3", gen_normal(shape, 1e-3, base_seed + 1)
    yield "n0.02", gen_normal(shape, 0.02, base_seed + 2)
    yield "n1", gen_normal(shape, 1.0, base_seed + 3)
    yield "n30", gen_normal(shape, 30.0, base_seed + 4)
    yield "cliffs", gen_cliffs(shape, base_seed + 5)
    yield "tiny", gen_tiny(shape, base_seed + 6)
    yield "zeroblk", gen_zero_blocks(shape, 1.0, base_seed + 7)
    yield "allzero", torch.zeros(shape)
    if len(shape) >= 2:
        yield "softmax", gen_softmax(shape, base_seed + 8)


def call_ref(Q, name, kw, x, block_size, skip):
    kw = {k: (None if v == "NA->None" else v) for k, v in kw.items()}
    fn = Q.QUANTIZER_MAP[name]
    if name in BLOCKED:
        return fn(x, block_size=list(block_size), skip_first_dim=skip, **kw)
    return fn(x, **kw)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_loader.load_quantize()
    Q = ref.quantizers
    arrays, manifest = {}, []
    idx = 0
    for li, (ltag, shape, skip) in enumerate(layouts()):
        for itag, x in inputs_for(shape, 1000 * (li + 1)):
            for fi, (name, kw) in enumerate(FORMATS):
                if name in BLOCKED:
                    bss = [[1, 16], [16]] if len(shape) > 1 else [[16], [1, 16]]
                    # general 2-D / odd block shapes on a subset to bound fixture size
                    if itag in ("n1", "zeroblk", "cliffs") and fi in (0, 5, 8):
                        bss = bss + [[2, 16], [16, 16], [4], [3, 5], [1, 32], [1, 8], [1, 64]]
                else:
                    bss = [None]
                    if ltag not in ("act2d_7x40", "act3d_2x16x64"):
                        continue
                for bs in bss:
                    y = call_ref(Q, name, kw, x, bs, skip)
                    key = f"c{idx:04d}"
                    xin = f"x_{ltag}_{itag}"
                    if xin not in arrays:
                        arrays[xin] = bits(x)
                    arrays[key] = bits(y)
                    manifest.append(dict(key=key, x=xin, fmt=name, kwargs=kw, block_size=bs, skip_first_dim=skip,
                                         layout=ltag, input=itag))
                    idx += 1
    # extreme magnitudes (saturation, denormals, exponent clamps) — bfp default bias & narrow exponents
    x = gen_extreme((4, 64), 77)
    arrays["x_extreme"] = bits(x)
    for name, kw in FORMATS:
        bs = [1, 16] if name in BLOCKED else None
        y = call_ref(Q, name, kw, x, bs, True if name in BLOCKED else False)
        key = f"c{idx:04d}"
        arrays[key] = bits(y)
        manifest.append(dict(key=key, x="x_extreme", fmt=name, kwargs=kw, block_size=bs, skip_first_dim=True,
                             layout="act2d_4x64", input="extreme"))
        idx += 1
    # non-contiguous (transposed view) input, the kT case of bmm_0 (modeling_opt.py:246)
    base = gen_normal((3, 48, 8), 1.0, 4242)
    xt = base.transpose(1, 2)                       # [3, 8, 48] view, blocks along the strided dim
    arrays["x_kT_base"] = bits(base)
    for name, kw in FORMATS[:1] + FORMATS[5:6] + FORMATS[8:9]:
        y = call_ref(Q, name, kw, xt, [1, 16], True)
        key = f"c{idx:04d}"
        arrays[key] = bits(y)
        manifest.append(dict(key=key, x="x_kT_base", transpose=[1, 2], fmt=name, kwargs=kw, block_size=[1, 16],
                             skip_first_dim=True, layout="kT_3x8x48", input="n1"))
        idx += 1
    np.savez_compressed(os.path.join(GOLD, "quantizers.npz"), **arrays)
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(dict(reference_commit="740bf4834cc91c9aa109cc86d2537d254f356137", torch=torch.__version__,
                       cases=manifest), f, indent=0)
    print(f"quantizer cases: {idx}")

    # ---- hashed larger cases ------------------------------------------------
    hashed = []
    big = [("act2d_256x4096", (256, 4096), True, 1.0), ("w2d_512x1024", (512, 1024), False, 0.02),
           ("act3d_4x256x1024", (4, 256, 1024), True, 1.0), ("probs_4x256x256", (4, 256, 256), True, None)]
    for tag, shape, skip, sigma in big:
        seed = 9000 + len(hashed)
        x = gen_softmax(shape, seed) if sigma is None else gen_normal(shape, sigma, seed)
        for name, kw in FORMATS[:2] + FORMATS[5:7] + FORMATS[8:10] + FORMATS[11:12]:
            y = call_ref(Q, name, kw, x, [1, 16], skip)
            hashed.append(dict(tag=tag, shape=list(shape), skip_first_dim=skip, sigma=sigma, seed=seed, fmt=name,
                               kwargs=kw, block_size=[1, 16] if name in BLOCKED else None,
                               sha256=hashlib.sha256(bits(y).tobytes()).hexdigest()))
    with open(os.path.join(GOLD, "hashed.json"), "w") as f:
        json.dump(hashed, f, indent=0)
    print(f"hashed cases: {len(hashed)}")

    # ---- consumers: Linear + matmul/bmm --------------------------------------
    import toml

    carr, cman = {}, []
    cfgs = {}
    for fn in sorted(os.listdir(REF_CFG_DIR)):
        cfgs[fn] = toml.load(os.path.join(REF_CFG_DIR, fn))
    from copy import deepcopy

    def na(d):
        return {k: (None if v == "NA" else v) for k, v in d.items()}

    lin_cfgs = {
        "bfp_6bit": na(cfgs["bfp_6bit.toml"]["default"]),
        "bfp_4bit": na(cfgs["bfp_4bit.toml"]["default"]),
        "block_minifloat": na(cfgs["block_minifloat.toml"]["default"]),
        "block_log": na(cfgs["block_log.toml"]["default"]),
        "minifloat_denorm": na(cfgs["minifloat_denorm.toml"]["default"]),
        "bypass": na(cfgs["bypass.toml"]["default"]),
    }
    ci = 0
    for cname, cfg in lin_cfgs.items():
        for xshape in [(5, 80), (2, 7, 80)]:
            torch.manual_seed(ci)
            scale = 4.0 if cname == "block_minifloat" else 0.05
            w = gen_normal((48, 80), scale, 500 + ci)
            b = gen_normal((48,), scale, 600 + ci)
            x = gen_normal(xshape, 1.0, 700 + ci)
            cls = ref.modules.QUANTIZED_MODULE_MAP["linear"][cfg["name"]] if not cfg.get("bypass") else \
                ref.modules.QUANTIZED_MODULE_MAP["linear"]["block_fp"]
            lin = cls(80, 48, bias=True, config=deepcopy(cfg))
            with torch.no_grad():
                lin.weight.copy_(w)
                lin.bias.copy_(b)
                y = lin(x)
            key = f"lin{ci:02d}"
            carr[key + "_x"], carr[key + "_w"], carr[key + "_b"] = bits(x), bits(w), bits(b)
            carr[key + "_y"] = bits(y)
            carr[key + "_wq"], carr[key + "_bq"] = bits(lin.weight.data), bits(lin.bias.data)
            cman.append(dict(key=key, op="linear", config_name=cname, config=cfg))
            ci += 1
    for cname in ["bfp_6bit", "block_minifloat", "block_log", "minifloat_denorm"]:
        cfg = lin_cfgs[cname]
        for style, xs, ys, tr in [("bmm", (6, 20, 32), (6, 24, 32), True), ("bmm", (6, 20, 48), (6, 48, 32), False),
                                  ("matmul", (2, 3, 20, 32), (2, 3, 24, 32), True), ("matmul", (20, 48), (48, 32), False)]:
            x = gen_normal(xs, 1.0, 800 + ci)
            yb = gen_normal(ys, 1.0, 900 + ci)
            yv = yb.transpose(-1, -2) if tr else yb
            f = ref.functions.QUANTIZED_FUNC_MAP[style][cfg["name"]]
            out = f(x, yv, config=deepcopy(cfg))
            key = f"mm{ci:02d}"
            carr[key + "_x"], carr[key + "_y"], carr[key + "_o"] = bits(x), bits(yb), bits(out)
            cman.append(dict(key=key, op=style, config_name=cname, config=cfg, y_transposed=tr))
            ci += 1
    np.savez_compressed(os.path.join(GOLD, "consumers.npz"), **carr)
    with open(os.path.join(GOLD, "consumers.json"), "w") as f:
        json.dump(cman, f, indent=0)
    print(f"consumer cases: {ci}")

    # ---- config surface ------------------------------------------------------
    models = ref_loader.load_models()
    out = {"raw": cfgs, "node": [], "opt": {}, "llama": {}}
    for fn, raw in cfgs.items():
        d = na(raw["default"])
        for op in ("linear", "matmul", "bmm", "rotary_positional_encoding"):
            try:
                out["node"].append(dict(file=fn, op=op, parsed=ref.parser.parse_node_config(deepcopy(d), op)))
            except Exception as e:  # e.g. KeyError on formats lacking a key
                out["node"].append(dict(file=fn, op=op, error=type(e).__name__))
        try:
            out["opt"][fn] = models.opt_qc.parse_opt_quantized_config(deepcopy(raw), 2)
        except Exception as e:
            out["opt"][fn] = {"error": type(e).__name__}
        try:
            out["llama"][fn] = models.llama_qc.parse_llama_quantized_config(deepcopy(raw), 2)
        except Exception as e:
            out["llama"][fn] = {"error": type(e).__name__}
    # a §4.4-style per-layer mixed precision config (search/opt_1.3b_sst2.toml:19-37 key paths)
    import random

    rnd = random.Random(0)
    mixed = {"default": deepcopy(cfgs["bfp_6bit.toml"]["default"])}
    for i in (0, 1):
        for path in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "self_attn.bmm_0",
                     "self_attn.bmm_1", "fc1", "fc2"):
            if i == 1 and path in ("fc1", "self_attn.bmm_1"):
                continue                              # unspecified → falls back to default
            node = deepcopy(cfgs["bfp_6bit.toml"]["default"])
            node["data_in_width"] = rnd.choice([6, 5, 4, 3])
            node["weight_width"] = rnd.choice([5, 4, 3, 2])
            node["bias_width"] = rnd.choice([5, 4, 3, 2])
            node["data_in_exponent_bias"] = node["weight_exponent_bias"] = node["bias_exponent_bias"] = "NA"
            cur = mixed.setdefault(f"model_layer_{i}", {})
            parts = path.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = node
    out["mixed_raw"] = mixed
    out["mixed_opt"] = models.opt_qc.parse_opt_quantized_config(deepcopy(mixed), 3)
    with open(os.path.join(GOLD, "configs.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("configs done")

    # ---- tiny-model forward goldens (reference modeling code, CPU fp32) -----------------------
    def sd_arrays(model):
        return {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}

    for tag, tomlname in [("opt_tiny_bfp6", "bfp_6bit.toml"), ("opt_tiny_bfp4", "bfp_4bit.toml"),
                          ("opt_tiny_mixed", None)]:
        torch.manual_seed(0)
        qc = deepcopy(cfgs[tomlname]) if tomlname else deepcopy(mixed)
        cfg = models.opt_cfg.OPTQuantizedConfig(hidden_size=64, num_hidden_layers=2, ffn_dim=128, num_attention_heads=4,
                                                vocab_size=512, max_position_embeddings=64, quant_config=qc)
        model = models.opt.OPTQuantizedForCausalLM(cfg).eval()
        arrs = sd_arrays(model)                      # weights BEFORE the in-place PTQ overwrite
        ids = torch.from_numpy(rs(5).randint(0, 512, size=(2, 64)).astype(np.int64))
        with torch.no_grad():
            o = model(input_ids=ids, labels=ids)
        arrs["input_ids"] = ids.numpy()
        arrs["logits"] = o.logits.numpy().copy()
        arrs["loss"] = np.array(float(o.loss))
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **arrs)
        print(tag, "loss", float(o.loss))

    # Llama: block_minifloat needs a scaled init (N(0,0.02) weights all quantise to 0)
    for tag, tomlname, init in [("llama_tiny_bmf8", "block_minifloat.toml", 1.5), ("llama_tiny_bl8", "block_log.toml", 0.02),
                                ("llama_tiny_bfp6", "bfp_6bit.toml", 0.02)]:
        torch.manual_seed(0)
        cfg = models.llama_cfg.LlamaQuantizedConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                                    num_attention_heads=4, vocab_size=512, max_position_embeddings=64,
                                                    initializer_range=init, quant_config=deepcopy(cfgs[tomlname]))
        model = models.llama.LlamaQuantizedForCausalLM(cfg).eval()
        arrs = {k: v for k, v in sd_arrays(model).items() if "rotary_emb" not in k}
        ids = torch.from_numpy(rs(6).randint(0, 512, size=(2, 64)).astype(np.int64))
        with torch.no_grad():
            o = model(input_ids=ids, labels=ids)
        arrs["input_ids"] = ids.numpy()
        arrs["logits"] = o.logits.numpy().copy()
        arrs["loss"] = np.array(float(o.loss))
        arrs["init"] = np.array(init)
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **arrs)
        print(tag, "loss", float(o.loss))


if __name__ == "__main__":
    main()
