#!/usr/bin/env python
"""Extracts, by line range, the function bodies of the reference's hot path that need nothing but a handful of class
members, into oracle/_ref/gen/*.inc (git-ignored, never committed: reference sources are not copied into the repo).
oracle/ref_pins_capi.cpp declares minimal stand-in classes with those members and #includes the fragments, so the
REFERENCE'S OWN statements are what gets compiled into oracle/_ref/libref_pins.so.  TEST INFRASTRUCTURE ONLY.

usage: python oracle/extract_ref.py /root/reference/planning_ddr_opt oracle/_ref/gen
"""
import sys
from pathlib import Path

FRAGMENTS = {
    # name: (file relative to planning_ddr_opt, first line, last line, sanity substring expected on the first line)
    "ref_sdf_esdf.inc": ("utils/plan_env/src/sdf_map.cpp", 618, 715, "void SDFmap::updateESDF2d()"),
    "ref_sdf_index.inc": ("utils/plan_env/src/sdf_map.cpp", 453, 472, "Eigen::Vector2d SDFmap::gridIndex2coordd(const Eigen::Vector2i &index)"),
    "ref_sdf_vecnum.inc": ("utils/plan_env/src/sdf_map.cpp", 525, 531, "int SDFmap::Index2Vectornum(const int &x, const int &y)"),
    "ref_sdf_lookup.inc": ("utils/plan_env/src/sdf_map.cpp", 739, 871, "inline double SDFmap::getDistance(const Eigen::Vector2i& id)"),
    "ref_sdf_isocc.inc": ("utils/plan_env/src/sdf_map.cpp", 942, 948, "bool SDFmap::isOccWithSafeDis(const Eigen::Vector2i &index, const double &safe_dis)"),
    "ref_minco_banded.inc": ("back_end/include/gcopter/minco.hpp", 43, 198, "class BandedSystem"),
    "ref_opt_tmaps.inc": ("back_end/src/optimizer.cpp", 573, 591, "template <typename EIGENVEC>"),
    "ref_opt_smoothl1.inc": ("back_end/src/optimizer.cpp", 1069, 1106, "inline void MSPlanner::positiveSmoothedL1"),
}


def main():
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    for name, (rel, a, b, expect) in FRAGMENTS.items():
        lines = (ref / rel).read_text().split("\n")
        frag = lines[a - 1:b]
        if expect not in frag[0]:
            raise SystemExit(f"{rel}:{a} does not start with {expect!r} (reference changed?): {frag[0]!r}")
        (out / name).write_text(f"// extracted from {rel}:{a}-{b} by oracle/extract_ref.py — do not edit, do not commit\n"
                                + "\n".join(frag) + "\n")
        print(f"{name}: {rel}:{a}-{b} ({len(frag)} lines)")


if __name__ == "__main__":
    main()
