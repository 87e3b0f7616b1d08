"""the raw-mesh FPS of preprocessing / inference (gen_utils.py:135-140, preprocess_data.py:55-56): one cloud of
n = 1e5..2e5 vertices -> 24 000.  Ours (auto shape and explicit bucket shapes) beside the verbatim reference kernel, indices bitwise."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from op_bench import ref_available, time_ms  # noqa: E402

from toothgroupnetwork_b200 import clouds, pointops  # noqa: E402


def main():
    rows = []
    for n in (50000, 100000, 200000):
        xyz = clouds.dental_arch(n, 7)[0].cuda().contiguous()
        off, noff = torch.tensor([n], dtype=torch.int32).cuda(), torch.tensor([24000], dtype=torch.int32).cuda()
        row = {"n": n, "m": 24000}
        got = pointops.fps_packed(xyz, off, noff, n, 24000)
        row["auto_ms"] = time_ms(lambda: pointops.fps_packed(xyz, off, noff, n, 24000), warm=1, reps=3)
        for mode in (-2, -26, -96, -112):
            try:
                alt = pointops.fps_packed(xyz, off, noff, n, 24000, mode)
                row[f"mode{mode}_ms"] = time_ms(lambda: pointops.fps_packed(xyz, off, noff, n, 24000, mode), warm=1, reps=3)
                row[f"mode{mode}_equal"] = bool(torch.equal(alt, got))
            except Exception as e:
                row[f"mode{mode}_error"] = repr(e)[:120]
        if ref_available():
            from oracle import ref_cuda
            ridx, _ = ref_cuda.furthestsampling(xyz, off, noff, n, 24000)
            row["idx_bitwise_vs_reference_kernel"] = bool(torch.equal(ridx, got))
            row["ref_kernel_ms"] = time_ms(lambda: ref_cuda.furthestsampling(xyz, off, noff, n, 24000), warm=0, reps=2)
            row["speedup"] = row["ref_kernel_ms"] / row["auto_ms"]
        row["us_per_iteration"] = row["auto_ms"] * 1e3 / 23999
        rows.append(row)
        print(json.dumps(row))
    with open(os.path.join(ROOT, "gpurun_out", "r2_raw_mesh_fps.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
