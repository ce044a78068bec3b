"""One-off source transformation: kernels named on the command line get pdl_enter() as their first statement and their
<<<...>>> launch sites become hnb::launch_pdl(...).  usage: pdl_convert.py file.cu kernel[:noenter] ..."""
import re, sys

def match_close(s, i, open_c, close_c):
    d = 0
    while i < len(s):
        if s[i] == open_c: d += 1
        elif s[i] == close_c:
            d -= 1
            if d == 0: return i
        i += 1
    raise ValueError("unbalanced")

def split_top(s):
    out, d, cur = [], 0, ""
    for ch in s:
        if ch in "([{": d += 1
        if ch in ")]}": d -= 1
        if ch == "," and d == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out

path, names = sys.argv[1], sys.argv[2:]
s = open(path).read()
for spec in names:
    name, _, flag = spec.partition(":")
    # 1. body
    if flag != "noenter":
        n = 0
        for m in list(re.finditer(r"__global__[^;{}]*?\b" + name + r"\s*\(", s))[::-1]:
            close = match_close(s, m.end() - 1, "(", ")")
            brace = s.index("{", close)
            assert s[close + 1:brace].strip() == "", (name, s[close + 1:brace])
            s = s[:brace + 1] + "\n  pdl_enter();" + s[brace + 1:]
            n += 1
        assert n >= 1, f"kernel {name} not found"
    # 2. launches
    n = 0
    pos = 0
    while True:
        m = re.search(r"\b" + name + r"\s*(<(?:[^<>;]|<[^<>;]*>)*>)?\s*<<<", s[pos:])
        if not m: break
        a, b = pos + m.start(), pos + m.end()
        kexpr = s[a:b - 3].strip()
        end = s.index(">>>", b)
        cfg = split_top(s[b:end])
        assert len(cfg) == 4, (name, cfg)
        par = s.index("(", end)
        assert s[end + 3:par].strip() == ""
        rep = f"hnb::launch_pdl({kexpr}, dim3({cfg[0]}), dim3({cfg[1]}), {cfg[2]}, {cfg[3]}, "
        s = s[:a] + rep + s[par + 1:]
        pos = a + len(rep)
        n += 1
    print(f"{path}: {name}: {n} launch site(s)")
open(path, "w").write(s)
