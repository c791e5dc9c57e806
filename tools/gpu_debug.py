"""Diagnostic dump: engine vs oracle, per tensor.  Run on the GPU box: python tools/gpu_debug.py [case ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from oracle import cvae_oracle as O
import parity_util as U

CASES = {
    "mm_z10_b48": (O.CVAEConfig(z_dim=10), 48, False),
    "mm_z10_b64_lab": (O.CVAEConfig(z_dim=10, num_classes=4), 64, True),
    "mm_z32_b24": (O.CVAEConfig(z_dim=32), 24, False),
    "uni_wave_b24": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50), 24, False),
    "uni_isi_b24_lab": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100, num_classes=4), 24, True),
    "mm_z10_b130": (O.CVAEConfig(z_dim=10), 130, False),
}


def main():
    names = sys.argv[1:] or list(CASES)
    verbose = os.environ.get("VERBOSE", "1") == "1"
    for name in names:
        cfg, B, lab = CASES[name]
        t0 = time.time()
        res, eng = U.run_train_case(cfg, B, lab, conv_path=int(os.environ.get("CONV_PATH", "0")))
        print("conv path in use:", eng.conv_path_in_use())
        print(f"=== {name}  ({time.time() - t0:.1f}s)  launches fwd+bwd {res['launches_fwd_bwd']} opt {res['launches_opt']}")
        print("loss eng", res["loss_eng"])
        print("loss f64", res["loss_f64"])
        print("loss rel (eng)", ["%.2e" % v for v in res["loss_rel"]], " (oracle f32)", ["%.2e" % v for v in res["loss_rel_f32"]])
        print("outputs  abs err eng / f32 / scale:", {k: tuple("%.2e" % x for x in v) for k, v in res["out_err"].items()})
        worst = sorted(res["tap_err"].items(), key=lambda kv: -kv[1][0])
        print("taps (rel max err eng, f32) worst 8:", [(k, "%.2e" % a, "%.2e" % b) for k, (a, b) in worst[:8]])
        if verbose:
            for k, (a, b) in res["tap_err"].items():
                if a > 1e-4:
                    print("   TAP", k, "%.3e" % a, "%.3e" % b)
        print("grad flat rel eng %.3e  oracle-f32 %.3e   |g| %.4f" % (res["grad_flat_rel"], res["grad_flat_rel_f32"], res["grad_global_norm"]))
        bad = []
        for n, (e, r, nn) in res["grad_err"].items():
            floor = 1e-6 * res["grad_global_norm"]
            if e > 3 * r + floor and e > 1e-4 * nn:
                bad.append((n, e, r, nn))
        print("grad tensors outside 3x oracle-f32 noise:", len(bad), "of", len(res["grad_err"]))
        for n, e, r, nn in bad[:60]:
            print("   GRAD %-55s err %.3e  f32err %.3e  norm %.3e  eng_norm %.3e cos %.4f" % ((n, e, r, nn) + res["grad_dbg"][n]))
        print("params with spurious grad:", res["no_grad_params"])
        print("running stats rel err %.2e | grad norm eng %.6f f64 %.6f | clip eng %.6f f64 %.6f" % (
            res["running_err"], res["grad_norm_eng"], res["grad_norm_f64"], res["clip_eng"], res["clip_f64"]))
        print("param abs err after AdamW %.3e | exp_avg rel %.3e | cls untouched %s" % (
            res["param_abs_err"], res["exp_avg_rel"], res.get("cls_emb_untouched")))
        ev, _ = U.run_eval_case(cfg, B, lab)
        print("eval abs err", {k: "%.2e" % v for k, v in ev["abs_err"].items()}, "embed", {k: "%.2e" % v for k, v in ev["emb_err"].items()},
              "zscore %.2e" % ev["zscore_err"], "loss rel", ["%.2e" % v for v in ev["loss_rel"]])
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
