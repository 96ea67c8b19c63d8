"""B200-native drop-in for the token hot path of karchkha/MelSpec_GPT_VQVAE.

    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import VectorQuantizer, LitVQVAE
    from melspec_gpt_vqvae_b200.transformer.minGPT import GPT, GPTClass, Lit_minGPT

All device work goes through libmgv.so (hand-written sm_100a CUDA behind the C ABI in
include/mgv.h).  No CPU fallback.
"""
__version__ = "0.1.0"
