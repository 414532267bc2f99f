"""Prints the end-to-end deviations of the CUDA path on the golden fixtures (both precisions) next to the tolerances."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import MODEL_CASES, load_golden
import test_gpu_parity as T
import hcunet_b200 as H

for name in MODEL_CASES:
    fx = load_golden(name)
    for precision in ("fp32", "mixed"):
        tol_out, tol_grad, agree_min = T.calibrated(T.TOL[precision], precision, fx["state_dict"], fx["kwargs"], fx["x"], fx["mask"],
                                                    fx["pwl"], fx["logits_train"], fx["grads"])
        m = T.build(fx, precision); m.train()
        x, mask, pwl = fx["x"].cuda(), fx["mask"].cuda(), fx["pwl"].cuda()
        logits = m(x)
        loss = H.cross_entropy(logits, mask, pwl, "pixel"); loss.backward()
        e = T.rel_l2(logits, fx["logits_train"])
        ge = max(T.rel_l2(p.grad, fx["grads"][k]) for k, p in m.named_parameters() if not T.is_dead_bias(k))
        print(f"{name:12s} {precision:5s} logits {e:.3e} (tol {tol_out:.3e})  worst grad {ge:.3e} (tol {tol_grad:.3e})")
