"""llm_mixed_q_b200.models — quantize/ (registries, kernels) and the quantized OPT / Llama / BERT module classes."""
