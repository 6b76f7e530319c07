"""Static SASS opcode census of the step kernels in the built library (cuobjdump -sass; no GPU needed).
Usage: python tools/sass_census.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rlao_b200", "libaoenv_b200.so")
WANT = ["shwfs_frame_kernel<6, 14>", "atm_phase_kernel", "gemm_tc_kernel<3, 128>", "gemm_tc_kernel<2, 256>", "atm_ring_kernel",
        "atm_rescan_kernel", "atm_gather_kernel", "shwfs_slopes_kernel<6>", "shwfs_detector_kernel", "dm_rows_kernel<12>",
        "pyr_rows_kernel<16, 18>", "pyr_image_kernel<16, 18>", "pyr_cols_kernel<16, 18>", "observe_kernel",
        "command_update_kernel"]
NOTE = {"FFMA2", "FADD2", "FMUL2", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "PREEXIT", "ACQBULK", "MUFU",
        "UTCATOMSWS", "UTMACCTL", "ELECT", "ATOMG", "RED", "STL", "LDL"}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            counts[kern][m.group(1)] += 1
    names = dict(zip(counts, subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()))
    print("# SASS opcode census of the step kernels (cuobjdump -sass of libaoenv_b200.so, sm_100a; tools/sass_census.py)\n")
    print("Static instruction counts per kernel body.  `FFMA2` / `FADD2` / `FMUL2` = packed FP32; `UTCHMMA` / `UTCBAR` / `LDTM` = "
          "tcgen05.mma / commit / tcgen05.ld; `UTMALDG` = TMA tensor load; `SYNCS` = mbarrier; `PREEXIT` / `ACQBULK` = "
          "griddepcontrol.launch_dependents / .wait; `MUFU` = SFU; `STL` / `LDL` = register spills.\n")
    for w in WANT:
        ks = [k for k, n in names.items() if w in n and "f64" not in n]
        if not ks:
            print(f"* `{w}`: not in the library")
            continue
        c = counts[ks[0]]
        top = ", ".join(f"{o} {n}" for o, n in c.most_common(10))
        nb = ", ".join(f"{o} {c[o]}" for o in sorted(NOTE) if c[o])
        print(f"* `{w}`: {sum(c.values())} instructions — {top}\n  - of note: {nb}")


if __name__ == "__main__":
    main()
