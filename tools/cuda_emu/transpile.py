"""Turn the library's .cu sources into C++ that g++ compiles against tools/cuda_emu/cuda_runtime.h: the three constructs that are
not C++ -- kernel launches `k<<<grid, block, smem, stream>>>(args)`, inline PTX and `extern __shared__` arrays -- are rewritten
textually; everything else is compiled as it stands.  Usage: transpile.py <in.cu> <out.cpp>"""
import re
import sys


def matching(s, i, open_ch, close_ch):
    """index of the bracket that closes the one at s[i] (string and character literals are skipped)"""
    depth, j = 0, i
    while j < len(s):
        c = s[j]
        if c == '"' or c == "'":
            q = c
            j += 1
            while s[j] != q:
                j += 2 if s[j] == "\\" else 1
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return j
        j += 1
    raise ValueError("unbalanced %s at %d" % (open_ch, i))


def split_top(s, sep=","):
    out, depth, cur, j = [], 0, "", 0
    while j < len(s):
        c = s[j]
        if c == '"':
            k = j + 1
            while s[k] != '"':
                k += 2 if s[k] == "\\" else 1
            cur += s[j:k + 1]
            j = k + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
        j += 1
    if cur.strip():
        out.append(cur.strip())
    return out


def strip_comments(s):
    """comments removed (string literals respected), line structure kept"""
    out, j = [], 0
    while j < len(s):
        c = s[j]
        if c == '"' or c == "'":
            q, k = c, j + 1
            while s[k] != q:
                k += 2 if s[k] == "\\" else 1
            out.append(s[j:k + 1])
            j = k + 1
        elif s.startswith("//", j):
            k = s.find("\n", j)
            j = len(s) if k < 0 else k
        elif s.startswith("/*", j):
            k = s.index("*/", j)
            out.append("\n" * s.count("\n", j, k))
            j = k + 2
        else:
            out.append(c)
            j += 1
    return "".join(out)


# ---- kernel launches ------------------------------------------------------------------------------------------------
def rewrite_launches(s):
    while True:
        i = s.find("<<<")
        if i < 0:
            return s
        j = i - 1                                              # the kernel's name (with template arguments) ends here
        while s[j].isspace():
            j -= 1
        end_name = j + 1
        if s[j] == ">":
            depth = 0
            while True:
                if s[j] == ">":
                    depth += 1
                elif s[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
            j -= 1
        while j >= 0 and (s[j].isalnum() or s[j] in "_:"):
            j -= 1
        name = s[j + 1:end_name]
        k = s.index(">>>", i)
        cfg = split_top(s[i + 3:k])
        a = s.index("(", k)
        b = matching(s, a, "(", ")")
        args = s[a + 1:b]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        # (arguments captured by value: a launch may be deferred, see EMU_DEFER_WORKERS in emu_runtime.cpp)
        new = 'emu::launch("%s", emu::d3(%s), emu::d3(%s), (size_t)(%s), [=]() { %s(%s); })' % (name, grid, block, smem, name, args)
        new = new.replace("\n", " ") + "\n" * s.count("\n", j + 1, b + 1)     # (line numbers stay those of the .cu file)
        s = s[:j + 1] + new + s[b + 1:]


# ---- inline PTX -----------------------------------------------------------------------------------------------------
PTX = [
    (r"^mov\.b64 %0, \{%1, %2\};$", "{o0} = emu::pack2({i0}, {i1});"),
    (r"^fma\.rn\.f32x2 %0, %1, %2, %3;$", "{o0} = emu::fma2({i0}, {i1}, {i2});"),
    (r"^mov\.u32 %0, %%smid;$", "{o0} = 0u;"),
    (r"^mov\.u32 %0, %%warpid;$", "{o0} = threadIdx.x >> 5;"),
    (r"^dp4a\.u32\.s32 %0, %1, %2, %3;$", "{o0} = emu::dp4a_u32_s32({i0}, {i1}, {i2});"),
    (r"^ld\.global\.nc\.L1::no_allocate\.v4\.u32 \{%0,%1,%2,%3\}, \[%4\];$",
     "{{ const unsigned *emu_p = (const unsigned *)({i0}); {o0} = emu_p[0]; {o1} = emu_p[1]; {o2} = emu_p[2]; {o3} = emu_p[3]; }}"),
    (r"^ld\.shared\.v4\.u32 \{%0,%1,%2,%3\}, \[%4\];$",
     "{{ const unsigned *emu_p = (const unsigned *)emu::from_shared({i0}); {o0} = emu_p[0]; {o1} = emu_p[1]; {o2} = emu_p[2]; {o3} = emu_p[3]; }}"),
    (r"^ld\.shared\.v2\.u32 \{%0,%1\}, \[%2\];$",
     "{{ const unsigned *emu_p = (const unsigned *)emu::from_shared({i0}); {o0} = emu_p[0]; {o1} = emu_p[1]; }}"),
    (r"^st\.shared\.v2\.u32 \[%0\], \{%1,%2\};$",
     "{{ unsigned *emu_p = (unsigned *)emu::from_shared({i0}); emu_p[0] = {i1}; emu_p[1] = {i2}; }}"),
    (r"^ld\.shared\.u32 %0, \[%1\];$", "{o0} = *(const unsigned *)emu::from_shared({i0});"),
    (r"^st\.shared\.u32 \[%0\], %1;$", "*(unsigned *)emu::from_shared({i0}) = {i1};"),
    (r"^$", ""),
]


def rewrite_asm(s):
    out, pos = [], 0
    for m in re.finditer(r"\basm\b(\s+volatile)?\s*\(", s):
        if m.start() < pos:
            continue
        a = m.end() - 1
        b = matching(s, a, "(", ")")
        semi = s.index(";", b)
        raw, depth, cur, j = [], 0, "", 0                      # the sections between top-level colons (empty ones kept)
        body = s[a + 1:b]
        while j < len(body):
            c = body[j]
            if c == '"':
                k = j + 1
                while body[k] != '"':
                    k += 2 if body[k] == "\\" else 1
                cur += body[j:k + 1]
                j = k + 1
                continue
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            if c == ":" and depth == 0:
                raw.append(cur)
                cur = ""
            else:
                cur += c
            j += 1
        raw.append(cur)
        template = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', raw[0])).strip()

        def operands(sec):
            ops = []
            for item in split_top(sec):
                mm = re.match(r'^"[^"]*"\s*\((.*)\)$', item, re.S)
                if mm:
                    ops.append(mm.group(1).strip())
            return ops
        outs = operands(raw[1]) if len(raw) > 1 else []
        ins = operands(raw[2]) if len(raw) > 2 else []
        repl = None
        for pat, form in PTX:
            if re.match(pat, template):
                names = {"o%d" % k: v for k, v in enumerate(outs)}
                names.update({"i%d" % k: v for k, v in enumerate(ins)})
                # "+r" operands are listed with the outputs and are not numbered among the inputs of these forms
                repl = form.format(**names) if form else ""
                break
        if repl is None:
            repl = 'emu::unsupported_asm("%s");' % template.replace("\\", "\\\\").replace('"', '\\"')[:200]
        out.append(s[pos:m.start()])
        out.append(repl.replace("\n", " ") + "\n" * s.count("\n", m.start(), semi + 1))
        pos = semi + 1
    out.append(s[pos:])
    return "".join(out)


# ---- dynamic shared memory ------------------------------------------------------------------------------------------
def rewrite_extern_shared(s):
    def sub(m):
        decl = re.sub(r"__align__\s*\([^)]*\)", "", m.group(1)).strip()
        return "%s *%s = (%s *)emu::dynamic_smem();" % (decl, m.group(2), decl)
    return re.sub(r"extern\s+__shared__\s+((?:__align__\s*\([^)]*\)\s*)?[A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;", sub, s)


def transpile(text):
    text = strip_comments(text)
    text = rewrite_extern_shared(text)
    text = rewrite_asm(text)
    text = rewrite_launches(text)
    return '#include "cuda_runtime.h"  // (kept on line 1: the line numbers below are those of the original file)\n' + text.split("\n", 1)[1] if text.startswith("//") or text.startswith("\n") else '#include "cuda_runtime.h"\n' + text


if __name__ == "__main__":
    with open(sys.argv[1]) as f:
        src = f.read()
    with open(sys.argv[2], "w") as f:
        f.write(transpile(src))
