"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

Copies the hot-path modules of the LIVE reference (``/root/reference/src``, read-only) into the git-ignored
``baseline/_ref/src`` so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference ITSELF on the
GPU box's host cores (the box has no ``/root/reference``; ``baseline/_ref`` travels with the ``gpurun`` snapshot like the built
``.so``).  Nothing under ``baseline/_ref`` is tracked by git; no reference source enters the repository's history.

    python oracle/vendor_reference.py            # run in the build container (``__graft_entry__.build()`` does it too)
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("IDB200_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

# the modules `src.sample.sample_generate` / `src.train.train_interp_levels` import for the maze path;
# the video branch (wan*, didemo, lsmdc, sinkhorn, latent_*) is not needed and not copied
FILES = [
    "src/__init__.py",
    "src/corruptions/__init__.py", "src/corruptions/keyframes.py",
    "src/data/__init__.py", "src/data/dataset.py", "src/data/maze.py", "src/data/astar.py", "src/data/trajectories.py",
    "src/diffusion/__init__.py", "src/diffusion/ddpm.py", "src/diffusion/schedules.py",
    "src/eval/__init__.py", "src/eval/metrics.py", "src/eval/visualize.py",
    "src/models/__init__.py", "src/models/transformer.py", "src/models/encoders.py", "src/models/denoiser_keypoints.py",
    "src/models/denoiser_interp_levels.py", "src/models/denoiser_interp_levels_causal.py", "src/models/keypoint_selector.py",
    "src/models/segment_cost.py",
    "src/selection/__init__.py", "src/selection/epiplexity_dp.py",
    "src/sample/__init__.py", "src/sample/sample_generate.py",
    "src/train/__init__.py", "src/train/train_interp_levels.py",
    "src/utils/__init__.py", "src/utils/checkpoint.py", "src/utils/clamp.py", "src/utils/device.py", "src/utils/ema.py",
    "src/utils/logging.py", "src/utils/normalize.py", "src/utils/run_config.py", "src/utils/seed.py",
]


def vendor(verbose: bool = True) -> bool:
    """True when baseline/_ref is populated (copied now, or already there); False when the reference is absent."""
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.isdir(os.path.join(DST, "src"))
    n = 0
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):              # namespace package in the reference: make it a regular one
                open(dst, "a").close()
            continue
        shutil.copyfile(src, dst)
        n += 1
    if verbose:
        print(f"vendored {n} reference modules into {DST} (git-ignored)")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
