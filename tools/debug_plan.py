"""Run a model plan step by step with a device sync after each launch and report the first failing step.
Usage: python tools/debug_plan.py hubert|synth|rmvpe simt|tc [seconds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicovergen_b200 import tapgemm as tg  # noqa: E402
from aicovergen_b200.synthetic import make_hubert_state_dict, make_rmvpe_state_dict, make_rvc_checkpoint  # noqa: E402


def describe(st):
    if isinstance(st, tg.TapGemm):
        p = st.params
        return (f"TapGemm[{st.name}] N={p.N} Kc={p.Kc} taps={p.ntaps} OW={p.OW} OH={p.OH} OB={p.OB} box={p.BW}x{p.BH} "
                f"a_dim={list(p.a_dim)} a_stride={list(p.a_stride)} ldw={p.ldw} wstride={p.wstride} vec4={p.vec4} "
                f"o_s=({p.o_sb},{p.o_sh},{p.o_sw}) backend={st.backend} tc_ok={st.tc_supported()}")
    return repr(st)


def main():
    which, be = sys.argv[1], sys.argv[2]
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 1.37
    backend = tg.BACKEND_TC if be == "tc" else tg.BACKEND_SIMT
    if which == "hubert":
        from aicovergen_b200.hubert import HubertB200, _HubertPlan
        net = HubertB200(make_hubert_state_dict(), "cuda:0", backend)
        L = int(16000 * secs)
        plan = _HubertPlan(net, L, 12)
        plan.wav[:L].copy_(torch.randn(L, device="cuda") * 0.1)
    elif which == "rmvpe":
        from aicovergen_b200.rmvpe import RMVPEB200, _RmvpePlan
        net = RMVPEB200(make_rmvpe_state_dict(), device="cuda:0", backend=backend)
        plan = _RmvpePlan(net, int(16000 * secs))
        plan.audio.copy_(torch.randn(int(16000 * secs), device="cuda") * 0.1)
    else:
        raise SystemExit("unknown model")
    torch.cuda.synchronize()
    for i, st in enumerate(plan.steps):
        try:
            st()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"STEP {i} FAILED: {describe(st)}\n  {type(e).__name__}: {str(e)[:300]}")
            return 1
    print(f"all {len(plan.steps)} steps ok")
    return 0


if __name__ == "__main__":
    sys.exit(main())
