#!/usr/bin/env python
"""Do the tensor-bound match kernel and the ALU-bound prompt kernels of two volumes share the SMs?

Times, with CUDA events on one B200 (config 2 shapes): the match stage alone, the prompt stage alone, and both at the
same time on two streams (different volumes).  `concurrent` close to max(match, prompts) means the kernels are resident
on the same SMs; close to their sum means they exclude each other.  Experiment knobs are environment variables read by
the library (PSAM_TC_STAGES, PSAM_BW_CTAS, PSAM_NO_CARVEOUT)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protosam_b200 import synth  # noqa: E402
from protosam_b200.engine import CoarseVolumeEngine  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_chaos_mri"]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    vol = synth.make_volume(1234, Q=cfg["Q"], L=cfg["L"], C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    sup, fg, qry = (torch.from_numpy(a).to(dev) for a in (vol.sup, vol.fg, vol.qry))
    engs = [CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"]) for _ in range(2)]
    for e in engs:
        e.set_support(sup, fg)
    logits = engs[1].match(qry)
    engs[1].prompts_from_logits(logits)
    engs[0].match(qry)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(do_a, do_b):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        sa.wait_event(t0); sb.wait_event(t0)
        for _ in range(n):
            if do_a:
                with torch.cuda.stream(sa):
                    engs[0].match(qry)
            if do_b:
                with torch.cuda.stream(sb):
                    engs[1].prompts_from_logits(logits)
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n

    for _ in range(2):
        timed(True, True)
    res = {"match_ms": timed(True, False), "prompts_ms": timed(False, True), "concurrent_ms": timed(True, True),
           "env": {k: os.environ.get(k) for k in ("PSAM_TC_STAGES", "PSAM_BW_CTAS", "PSAM_NO_CARVEOUT")}}
    res["overlap"] = (res["match_ms"] + res["prompts_ms"] - res["concurrent_ms"]) / min(res["match_ms"], res["prompts_ms"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
