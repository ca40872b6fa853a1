#!/usr/bin/env python
"""Static evidence from the built objects (no GPU needed): which SASS instructions the hot kernels contain — tcgen05.mma
shows up as UTCHMMA, tcgen05.ld / st as LDTM / STTM, tcgen05.commit as UTCBAR, cp.async.bulk (shared -> global) as UBLKCP,
mbarrier waits as SYNCS — plus registers / spills per kernel.  Writes profiles/r02_sass_evidence.md.

    python profiles/sass_evidence.py
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "q_learning_with_hjb_b200", "build", "obj")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU", "FFMA", "LDC", "STG", "REDG", "ATOMG"]
KERNELS = [   # (object, substring of the mangled name, label)
    ("vhjb_tc_quad10d.o", "14vhjb_tc_kernelINS_10Quad10DSysILb0EEELi0ELi0ELi0ELb1ELi0ELb0", "vhjb_tc_kernel<Quad10D, relu, GRAD> (C5 gradient)"),
    ("vhjb_tc_quad10d.o", "14vhjb_tc_kernelINS_10Quad10DSysILb0EEELi0ELi0ELi0ELb1ELi0ELb1", "vhjb_tc_kernel<Quad10D, relu, GRAD, STREAM>"),
    ("vhjb_tc_quad10d.o", "24vhjb_tc_residual2_kernelINS_10Quad10DSysILb0EEELi0", "vhjb_tc_residual2_kernel<Quad10D, relu> (C5 residual: states on lanes, TS-mode MMAs)"),
    ("vhjb_tc_quad10d.o", "23vhjb_tc_residual_kernelINS_10Quad10DSysILb0EEELi0", "vhjb_tc_residual_kernel<Quad10D, relu> (round 1's residual kernel, HJB_VHJB_RESIDUAL=v1)"),
    ("vhjb_tc_linear21_sin.o", "14vhjb_tc_kernelINS_9LinearSysILi2ELi1ELb0EEELi2ELi1ELi1ELb1ELi0ELb0", "vhjb_tc_kernel<Linear21, sin, min-time, GRAD> (C2)"),
    ("rollout_quad2d_fb.o", "Quad2DSysILb1EEENS_11FeedbackCtlILb1EEELi0ELb0ELi3ELb0", "rollout_kernel<Quad2D fast, hover LQR, Euler, final + unit cost> (C4)"),
    ("rollout_acrobot_es.o", "AcrobotSysILb1EEENS_12AcrobotESCtlELi0ELb0ELi3ELb0", "rollout_kernel<Acrobot fast, energy shaping, Euler, final + unit cost> (C3)"),
    ("rollout_quad10d_fb.o", "Quad10DSysILb1EEENS_11FeedbackCtlILb1EEELi0ELb1ELi0ELb0", "rollout_kernel<Quad10D fast, hover LQR, Euler, RECORD> (TMA bulk stores)"),
]


def sass(obj, needle):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn = next((m for m in re.findall(r"Function : (\S+)", names) if needle in m), None)
    if fn is None:
        return None, None, ""
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    usage = ""
    lines = res.splitlines()
    for i, line in enumerate(lines):
        if fn in line and i + 1 < len(lines):
            usage = lines[i + 1].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt, flags=re.M)
    return fn, ops, usage


def main():
    out = ["# r02 — SASS evidence (static, from the built objects)", "",
           "`python profiles/sass_evidence.py`; mnemonics per `/opt/skills/guides/B200_PROFILING.md`: `tcgen05.mma` → `UTCHMMA`, "
           "`tcgen05.ld/st` → `LDTM/STTM`, `tcgen05.commit` → `UTCBAR`, `cp.async.bulk` (shared → global) → `UBLKCP`, mbarrier → `SYNCS`.", ""]
    for obj, needle, label in KERNELS:
        fn, ops, usage = sass(os.path.join(OBJ, obj), needle)
        out.append(f"## {label}")
        if fn is None:
            out += ["(not found in the built objects)", ""]
            continue
        counts = {w: sum(1 for o in ops if o.split(".")[0] == w) for w in WATCH}
        out += [f"`{fn[:110]}`", "", f"{len(ops)} SASS instructions; {usage}", "",
                "| " + " | ".join(WATCH) + " |", "|" + "---:|" * len(WATCH),
                "| " + " | ".join(str(counts[w]) for w in WATCH) + " |", ""]
    path = os.path.join(ROOT, "profiles", "r02_sass_evidence.md")
    with open(path, "w") as fh:
        fh.write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
