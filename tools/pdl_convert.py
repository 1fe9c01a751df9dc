"""One-off source transformation (kept for the record): turn every `kernel<<<grid, block, smem, stream>>>(args);` of
coper_b200/csrc into `launch_pdl(kernel, grid, block, smem, stream, args);` and make `pdl_enter();` the first
statement of every __global__ function, so that consecutive kernels of a stream are linked by programmatic dependent
launch (common.cuh).  Idempotent: files that already use launch_pdl are left alone.

    python tools/pdl_convert.py [--check]
"""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent / "coper_b200" / "csrc"
SKIP_ENTER = {"umma_gemm_kernel", "bce_dq_fused_kernel"}     # these place trigger / wait themselves (after the prologue)


def match_close(s, i, open_ch, close_ch):
    """index of the bracket closing the one at s[i]"""
    depth = 0
    while i < len(s):
        if s[i] == open_ch:
            depth += 1
        elif s[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def kernel_expr_start(s, end):
    """start index of the kernel expression that ends right before s[end:end+3] == '<<<' (identifier with optional
    template arguments)"""
    i = end
    if s[i - 1] == ">":                      # template arguments
        depth = 0
        while True:
            i -= 1
            if s[i] == ">":
                depth += 1
            elif s[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while i > 0 and (s[i - 1].isalnum() or s[i - 1] in "_:"):
        i -= 1
    return i


def convert_launches(src):
    out, pos, n = "", 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            break
        line_start = src.rfind("\n", 0, k) + 1
        if "//" in src[line_start:k]:          # inside a comment
            out += src[pos:k + 3]
            pos = k + 3
            continue
        ks = kernel_expr_start(src, k)
        cfg_end = src.find(">>>", k)
        cfg = split_top(src[k + 3:cfg_end])
        assert len(cfg) == 4, (cfg, src[line_start:cfg_end])
        a0 = cfg_end + 3
        assert src[a0] == "(", src[line_start:a0 + 20]
        a1 = match_close(src, a0, "(", ")")
        assert src[a1 + 1] == ";", src[line_start:a1 + 2]
        args = " ".join(src[a0 + 1:a1].split())
        kernel = src[ks:k]
        call = "launch_pdl(%s, %s, %s, %s, %s, %s);" % (kernel, cfg[0], cfg[1], cfg[2], cfg[3], args)
        # re-wrap at 120 columns with the indentation of the statement
        indent = " " * (ks - line_start) if src[line_start:ks].strip() == "" else ""
        if indent and len(indent) + len(call) > 118:
            words, lines, cur = call.split(" "), [], indent
            for w in words:
                if len(cur) + len(w) + 1 > 118 and cur.strip():
                    lines.append(cur.rstrip())
                    cur = indent + "           " + w + " "
                else:
                    cur += w + " "
            lines.append(cur.rstrip())
            call = "\n".join(lines).lstrip()
        out += src[pos:ks] + call
        pos = a1 + 2
        n += 1
    return out + src[pos:], n


GLOBAL_RE = re.compile(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(")


def add_enter(src):
    out, pos, n = "", 0, 0
    for m in GLOBAL_RE.finditer(src):
        name = m.group(1)
        p0 = m.end() - 1
        p1 = match_close(src, p0, "(", ")")
        b = src.find("{", p1)
        if src[p1 + 1:b].strip() != "":
            continue                           # declaration or something unexpected
        if name in SKIP_ENTER:
            continue
        body_next = src[b + 1:b + 60]
        if "pdl_enter();" in body_next:
            continue
        out += src[pos:b + 1] + "\n  pdl_enter();"
        pos = b + 1
        n += 1
    return out + src[pos:], n


def main():
    check = "--check" in sys.argv
    for f in sorted(list(ROOT.glob("*.cu")) + list(ROOT.glob("*.cuh"))):
        src = f.read_text()
        new, n_l = convert_launches(src)
        new, n_e = add_enter(new)
        if new != src:
            print("%-20s launches %2d  kernels %2d" % (f.name, n_l, n_e))
            if not check:
                f.write_text(new)


if __name__ == "__main__":
    main()
