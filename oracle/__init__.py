"""CPU oracle for the Qwen3-TTS decode hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  The product path (`qwen3_tts_rs_b200`) never imports `oracle` and fails
loudly when its CUDA library is missing.

What it is: a torch/numpy restatement, op for op, of the reference's algorithm for the
path named in BASELINE.json `north_star` -- each function cites the reference file:line it
follows.  Two precision modes:
  * F32   : the reference's CPU path (src/lib.rs:1436-1442 -> F32 on CPU)
  * BF16  : the reference's CUDA path; every candle op writes a bf16 tensor, so values are
            rounded to bf16 after every op the reference executes as a separate candle op
            (f32 accumulation inside matmul / rms_norm / softmax, as candle + cuBLAS do).

PARITY STATUS: **parity unpinned** for the model forward passes.  The reference cannot be
built here (no cargo/rustc, candle not vendored), its golden files (test_data/) are
git-ignored upstream and absent, and no Qwen3-TTS weights exist in this environment.
What IS pinned (tests/test_oracle_*.py):
  * every weight-free known answer in the reference's own unit tests (sampler algebra,
    suppression mask, penalty mask, codes_to_tensor layout, causal-conv causality/length,
    trans-conv lengths, SnakeBeta alpha=beta=0, fused==sequential RMSNorm, vocoder stage
    shapes, upsample total 1920) -- SURVEY.md §8(c);
  * the reference's one native kernel, kernels/fused_residual_rmsnorm.cu, IS compilable:
    oracle/build_ref.py compiles it from where it lies into oracle/_ref/ and the GPU tests
    compare both this oracle and the product kernel against it bit for bit;
  * block semantics of the vocoder cross-checked against transformers' qwen3_omni_moe
    Code2Wav modules (same model family) in tests/test_oracle_vocoder.py.
Third-party arithmetic the reference delegates to candle 0.9 / cuBLAS / flash-attn (not in
/root/reference) is restated from those libraries' documented algorithms; the assumptions
are listed in DESIGN.md §Oracle.
"""
