"""
Analytic cost model of a quantized layer: parameter / activation counts, storage bits and FLOPs — the numbers behind the
reference's "memory density" and "arithmetic density" objectives.

Mirror of reference models/quantize/quantized_layer_profiler.py:10-177 (`profile_linear_layer`, `profile_matmul_layer`,
`update_profile`) with the same argument names, dictionary keys, integer types and failure mode (`ValueError("Unknown
quant_arith: ...")` for anything but bypass / integer / block_fp).  Pure host arithmetic: nothing here touches a tensor.
The statistic-hook half of that file (`register_a_stat_hook`, :180-206) belongs to the statistic profiler, which is out of scope.

Storage model (reference :10-30): a plain or integer tensor costs numel * width bits; a block_fp tensor is padded up to whole
blocks, every element costs `width` bits and every block one shared exponent of `exponent_width` bits.
"""
from __future__ import annotations

import numpy as np

_KEYS = ("num_params", "num_acts", "param_bits", "act_bits", "flops")


def compute_tensor_bits_fp(tensor_shape, width: int):
    return np.prod(np.asarray(tensor_shape)) * width


compute_tensor_bits_integer = compute_tensor_bits_fp


def compute_tensor_bits_block_fp(tensor_shape, width: int, exponent_width: int, block_size):
    shape, block = np.asarray(tensor_shape), np.asarray(block_size)
    if shape.size > block.size:                       # leading dims are not blocked
        block = np.append([1] * (shape.size - block.size), block)
    elif shape.size < block.size:                     # right-aligned, like the quantizer (quantizers/utils.py:42-67)
        block = block[-shape.ndim:]
    num_blocks = np.prod(np.ceil(shape / block))
    return num_blocks * np.prod(block) * width + num_blocks * exponent_width


def _operand_bits(quant_config: dict, prefix: str, shape, bypass: bool):
    """Storage bits of one operand (`prefix` in data_in / weight / bias) of `shape` under `quant_config`."""
    if bypass:
        return compute_tensor_bits_fp(shape, 32)
    arith = quant_config["name"]
    if arith == "integer":
        return compute_tensor_bits_integer(shape, quant_config[f"{prefix}_width"])
    if arith == "block_fp":
        return compute_tensor_bits_block_fp(shape, quant_config[f"{prefix}_width"], quant_config[f"{prefix}_exponent_width"],
                                            np.array(quant_config[f"{prefix}_block_size"]))
    raise ValueError(f"Unknown quant_arith: {arith}")


def _as_profile(num_params, num_acts, param_bits, act_bits, flops) -> dict:
    vals = (num_params, num_acts, param_bits, act_bits, flops)
    return {k: np.rint(v).astype(np.int64) for k, v in zip(_KEYS, vals)}


def profile_linear_layer(quant_config: dict, in_features: int, out_features: int, bias: bool, batch_size: int) -> dict:
    """x [batch_size, in_features] @ w [in_features, out_features] (+ b [out_features]).  Reference :33-118."""
    # like the reference, the widths are looked up before the bypass test: a config without them raises KeyError either way
    quant_config["weight_width"], quant_config["data_in_width"]
    if bias:
        quant_config["bias_width"]
    bypass = bool(quant_config.get("bypass", False))
    w_shape, b_shape, x_shape = (in_features, out_features), (out_features,), (batch_size, in_features)
    param_bits = _operand_bits(quant_config, "weight", w_shape, bypass)
    if bias:
        param_bits = param_bits + _operand_bits(quant_config, "bias", b_shape, bypass)
    act_bits = _operand_bits(quant_config, "data_in", x_shape, bypass)
    num_params = in_features * out_features + (out_features if bias else 0)
    flops = batch_size * out_features * (2 * in_features - 1) + (batch_size * out_features if bias else 0)
    return _as_profile(num_params, batch_size * in_features, param_bits, act_bits, flops)


def profile_matmul_layer(quant_config: dict, data_in_0_size, data_in_1_size) -> dict:
    """x0 [M, K] @ x1 [K, N], both activations.  Reference :121-168, including its quirk: BOTH operands are charged
    `data_in_width` bits per element (x1 only takes its exponent width and block size from the `weight_*` keys).  The shapes are
    wrapped exactly as the reference wraps them (`np.array((size,))`, a 1 x 2 array), which only matters for how a block size
    with more entries than dimensions is truncated."""
    x0_shape, x1_shape = np.array((tuple(data_in_0_size),)), np.array((tuple(data_in_1_size),))
    num_acts = np.prod(x0_shape) + np.prod(x1_shape)
    width = quant_config["data_in_width"]
    if quant_config.get("bypass", False):
        act_bits = compute_tensor_bits_fp(x0_shape, 32) + compute_tensor_bits_fp(x1_shape, 32)
    elif quant_config["name"] == "integer":
        act_bits = compute_tensor_bits_integer(x0_shape, width) + compute_tensor_bits_integer(x1_shape, width)
    elif quant_config["name"] == "block_fp":
        act_bits = (compute_tensor_bits_block_fp(x0_shape, width, quant_config["data_in_exponent_width"],
                                                 np.array(quant_config["data_in_block_size"]))
                    + compute_tensor_bits_block_fp(x1_shape, width, quant_config["weight_exponent_width"],
                                                   np.array(quant_config["weight_block_size"])))
    else:
        raise ValueError(f"Unknown quant_arith: {quant_config['name']}")
    flops = data_in_0_size[0] * data_in_1_size[1] * (2 * data_in_0_size[1] - 1)
    return _as_profile(0, num_acts, 0, act_bits, flops)


def update_profile(profile: dict, delta: dict) -> dict:
    for k in _KEYS:
        profile[k] += delta[k]
    return profile


def empty_profile() -> dict:
    return {k: 0 for k in _KEYS}


def profile_transformer_layers(config, seq_len: int, layer_ops) -> dict:
    """Sum over `config.num_hidden_layers` layers of the ops `layer_ops(layer_quant_config)` yields: tuples
    ("linear", op_config, in_features, out_features, bias) or ("matmul", op_config, (M, K), (K, N)), the latter counted once
    per attention head by the caller.  Shared by the three per-model profilers (reference models/*/profiler_*.py)."""
    total = empty_profile()
    for i in range(config.num_hidden_layers):
        for op in layer_ops(config.quant_config[f"model_layer_{i}"]):
            if op[0] == "linear":
                _, qc, fin, fout, bias = op
                update_profile(total, profile_linear_layer(qc, in_features=fin, out_features=fout, bias=bias, batch_size=seq_len))
            else:
                _, qc, s0, s1 = op
                update_profile(total, profile_matmul_layer(qc, data_in_0_size=s0, data_in_1_size=s1))
    return total
