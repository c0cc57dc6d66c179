#!/usr/bin/env python
"""Builds the experiment variants of the library that scripts/gpu/knobs.sh and scripts/gpu/text_tile.sh compare
(tractor3d_b200/csrc/_build/libt3d_particles_<name>.so, selected at run time with T3D_LIBRARY=<path>).

    python scripts/build_variants.py [name ...]        (default: all)
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tractor3d_b200 import build  # noqa: E402

Q = lambda s: 'T3D_EXP_LD_QUAL="%s"' % s
VARIANTS = {
    # stores / loads of the update kernel's streaming accesses (profiles/r2_cache_hints.txt)
    "stplain": ["T3D_EXP_ST_PLAIN"], "stwt": ["T3D_EXP_ST_WT"],
    "ldcs": ["T3D_EXP_LD_CS"], "ldcg": ["T3D_EXP_LD_CG"],
    "ld256": [Q("ld.global.L2::256B")], "ld128": [Q("ld.global.L2::128B")], "ldna": [Q("ld.global.L1::no_allocate")],
    "ldna256": [Q("ld.global.L1::no_allocate.L2::256B")], "ldlu": [Q("ld.global.lu")],
    "cpa256": ['T3D_EXP_CPA_QUAL="cp.async.cg.shared.global.L2::256B"'], "cpa128": ['T3D_EXP_CPA_QUAL="cp.async.cg.shared.global.L2::128B"'],
    "bulkef": ['T3D_EXP_BULK_POLICY="evict_first"'], "bulkun": ['T3D_EXP_BULK_POLICY="evict_unchanged"'],
    # predecessors inspected per look-back round: 32 x this
    "look2": ["T3D_EXP_LOOK=2"], "look8": ["T3D_EXP_LOOK=8"],
    # characters per text-expansion CTA (the default is 128)
    "text256": ["T3D_EXP_TEXT_TILE=256"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    t = time.time()
    with ThreadPoolExecutor(6) as ex:
        for n, out in zip(names, ex.map(lambda n: build.build_variant(n, VARIANTS[n]), names)):
            print(n, out)
    print(f"{time.time() - t:.0f} s")
