"""One small invocation of every CUDA path of the library (the GPU tests' own functions at their smallest shapes), meant to
be run under compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize.py ops
The parity assertions of the tests stay active, so a pass means: no invalid access and unchanged results."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch   # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(name, fn, *a):
    fn(*a)
    torch.cuda.synchronize()
    print("ok", name, a, flush=True)


import test_gpu_ops as T   # noqa: E402
run("gemm", T.test_gemm, 128, 256, 64, 0, 0, False, False, 0)
run("gemm", T.test_gemm, 300, 768, 768, 0, 0, True, False, 0)          # 2-CTA kernel? small-M path: 128-wide tiles
run("gemm", T.test_gemm, 257, 1024, 768, 1, 1, True, False, 0)
run("gemm", T.test_gemm, 513, 768, 1024, 0, 0, True, True, 0)
run("gemm", T.test_gemm, 200, 128, 192, 0, 0, True, False, 7)
run("gemm", T.test_gemm, 128 * 170, 2304, 768, 1, 0, True, False, 0)   # CTA-pair kernel, fp16 TMA-store epilogue
run("attention", T.test_attention, 2, 30, None, 0)
run("attention", T.test_attention, 3, 100, "rand", 1)
run("attention", T.test_attention, 1, 300, None, 0)
run("attention", T.test_attention, 2, 257, "rand", 0)
run("attention", T.test_attention, 2, 1000, "blocks", 1)
run("layernorm", T.test_layernorm, 77, 0)
run("scheduler steps", T.test_ddpm_and_pndm_step_kernels)
if what != "ops-no-res1":
    run("gemm", T.test_gemm, 128 * 170 + 5, 768, 1024, 0, 0, True, True, 0)   # CTA-pair kernel, TMA-staged residual epilogue
if what == "all":
    import test_gpu_cascade as C
    import test_gpu_compaction as K
    import test_gpu_denoisers as D
    import test_gpu_post as P
    import test_gpu_vae as V
    run("denoiser golden", D.test_golden, "surfpos", True)
    run("denoiser golden", D.test_golden, "edgez", False)
    run("compaction", K.test_compact_equals_dense_on_valid_tokens, *K.CASES[0])
    run("dedup surfaces", C.test_dedup_surfaces_bit_exact, 3, 7)
    run("dedup edges", C.test_dedup_edges_bit_exact, 2, 9, 30)
    run("short cascade", C.test_short_cascade_matches_oracle, True, "ddpm")
    run("surface decoder", V.test_surface_decoder, 5, 2)
    run("edge decoder", V.test_edge_decoder, 37, 16)
    run("encoders", V.test_encoders_match_oracle)
    run("post topology", P.test_topology_and_edges_match_reference, list(P.CASES)[0])
    run("post joint optimize", P.test_joint_optimize_matches_reference, list(P.CASES)[0])
print("SANITIZE_DONE", flush=True)
