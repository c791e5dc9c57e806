"""Gradient error table of the engine against the fp64 oracle, next to the reference's own fp32 error (the yardstick of
SURVEY.md section 8c): per case the worst per-tensor ratio err(engine, fp64) / err(reference fp32, fp64), the flat
relative L2 errors and the LeakyReLU sign flips.  Writes profiles/r02_grad_error_table.md.
    python tools/grad_table.py [out.md]         (on a B200; uses the oracle as the checker)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import cvae_oracle as O  # noqa: E402
import parity_util as U  # noqa: E402

CASES = {
    "mm_z10_b48": (O.CVAEConfig(z_dim=10), 48, False),
    "mm_z10_b64_labelled": (O.CVAEConfig(z_dim=10, num_classes=4), 64, True),
    "mm_z32_b64_labelled": (O.CVAEConfig(z_dim=32, num_classes=4), 64, True),
    "mm_z64_b64_labelled": (O.CVAEConfig(z_dim=64, num_classes=4), 64, True),
    "uni_wave_b24": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50), 24, False),
    "uni_isi_b24_labelled": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100, num_classes=4), 24, True),
    "mm_z10_b130": (O.CVAEConfig(z_dim=10), 130, False),
    "mm_z10_b256": (O.CVAEConfig(z_dim=10), 256, False),
    "mm_z10_b512": (O.CVAEConfig(z_dim=10), 512, False),
}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_grad_error_table.md")
    rows = []
    for name, (cfg, B, lab) in CASES.items():
        for path in (0, 1):
            res, _ = U.run_train_case(cfg, B, lab, conv_path=path)
            gn = res["grad_global_norm"]
            worst, worst_n, over2 = 0.0, "", 0
            for n, (e, r, nn) in res["grad_err"].items():
                ratio = e / max(r, 1e-6 * gn)  # the 1e-6 * ||g|| floor of SURVEY 8c
                if ratio > worst:
                    worst, worst_n = ratio, n
                over2 += ratio > 2.0
            wt, wt_n, ot = 0.0, "", 0
            for n, (e, r, nn) in res["grad_err_tf"].items():
                ratio = e / max(r, 1e-6 * gn)
                if ratio > wt:
                    wt, wt_n = ratio, n
                ot += ratio > 2.0
            rows.append((name, "tcgen05 pair" if path == 0 else "fp32 CUDA cores", max(res["loss_rel"]), res["grad_flat_rel"],
                         res["grad_flat_rel_f32"], worst, worst_n, over2, len(res["grad_err"]), res["flips_eng"], res["flips_f32"],
                         res["grad_flat_rel_tf"], res["grad_flat_rel_f32_tf"], wt, wt_n, ot, res["unforced_sites"][:3]))
            print(rows[-1], flush=True)
    with open(out, "w") as f:
        f.write("# Gradient error of the engine vs the fp64 oracle (teacher-forced, one step)\n\n"
                "`ratio` = worst per-tensor err(engine, fp64) / max(err(reference fp32, fp64), 1e-6 ||g||); flips = LeakyReLU "
                "inputs whose sign differs from the fp64 run (engine / reference fp32).\n\n"
                "`tf` columns: the same with the LeakyReLU branches of the fp64 run teacher-forced to the ones the compared run took.\n\n"
                "| case | conv path | max loss rel | flat rel-L2 (engine) | flat rel-L2 (ref fp32) | worst ratio | tensor | tensors > 2x | flips eng / f32 "
                "| tf flat (engine) | tf flat (ref fp32) | tf worst ratio | tf tensor | tf tensors > 2x |\n"
                "|---|---|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---|---:|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]} | {r[2]:.1e} | {r[3]:.1e} | {r[4]:.1e} | {r[5]:.2f} | {r[6]} | {r[7]} / {r[8]} | {r[9]} / {r[10]} "
                    f"| {r[11]:.1e} | {r[12]:.1e} | {r[13]:.2f} | {r[14]} | {r[15]} |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
