"""Where does the K=512 ratio contraction lose time?  Same shape, different epilogues."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from multimodal_b200 import _native
    M, N, K = 65536, 8192, 512
    ms = _native.contract_bench(M, N, K, "tf32", a_trans=False, b_trans=False, iters=5)
    print("%-40s %8.3f ms %7.1f TFLOP/s" % (sys.argv[1], ms, 2.0 * M * N * K / ms / 1e9), flush=True)
else:
    for name, env in [("store", {}), ("no store", {"KLNMF_BENCH_NOSTORE": "1"}),
                      ("ratio XT pair", {"KLNMF_BENCH_EPI": "ratio"}), ("ratio XT pair, KL only", {"KLNMF_BENCH_EPI": "ratio_kl"}),
                      ("ratio XT single CTA", {"KLNMF_BENCH_EPI": "ratio", "KLNMF_TC_CG": "1"}),
                      ("ratio XT single, KL only", {"KLNMF_BENCH_EPI": "ratio_kl", "KLNMF_TC_CG": "1"}),
                      ("ratio generic epilogue pair", {"KLNMF_BENCH_EPI": "ratio", "KLNMF_TC_NO_XT": "1"}),
                      ("ratio generic pair, KL only", {"KLNMF_BENCH_EPI": "ratio_kl", "KLNMF_TC_NO_XT": "1"})]:
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, name], env=e)
