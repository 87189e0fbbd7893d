"""Transcribes the reference's known-answer evaluator tests into JSON fixtures.

Sources (read here, in the build container; the GPU box has no /root/reference):
  fidget-core/src/eval/test/interval.rs   -> interval_known_answers.json
  fidget-core/src/eval/test/grad_slice.rs -> grad_known_answers.json

Each Rust test builds an expression with `ctx.<op>(..)`, evaluates it on
literal inputs and asserts literal outputs (and Choice traces).  This script
understands that small dialect; statements it cannot interpret are skipped and
counted.  The JSON is committed, the tests only read the JSON.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

NUM = r"-?(?:f32::\w+|std::f32::consts::\w+|\d+\.?\d*(?:e-?\d+)?(?:_?f32)?)"
CONSTS = {"f32::NAN": "nan", "f32::INFINITY": "inf", "-f32::INFINITY": "-inf",
          "std::f32::consts::PI": 3.14159265358979323846, "f32::EPSILON": 1.1920929e-07}


def num(tok):
    tok = tok.strip()
    if tok in CONSTS:
        return CONSTS[tok]
    if tok.startswith("-") and tok[1:] in CONSTS and isinstance(CONSTS[tok[1:]], float):
        return -CONSTS[tok[1:]]
    tok = tok.replace("_f32", "").replace("f32", "")
    return float(tok)


def interval(tok):
    tok = tok.strip()
    m = re.fullmatch(r"\[f32::NAN; 2\]", tok)
    if m:
        return ["nan", "nan"]
    m = re.fullmatch(r"\[\s*(%s)\s*,\s*(%s)\s*\]" % (NUM, NUM), tok)
    if not m:
        raise ValueError(tok)
    return [num(m.group(1)), num(m.group(2))]


def split_statements(body):
    out, depth, cur = [], 0, ""
    for ch in body:
        cur += ch
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        elif ch == ";" and depth == 0:
            out.append(" ".join(cur.split()))
            cur = ""
    return out


def parse_tests(path, prefix):
    src = open(path).read()
    tests = {}
    skipped = 0
    for fn in re.finditer(r"pub fn (%s\w+)\(\) \{" % prefix, src):
        name = fn.group(1)
        start = fn.end()
        # function body: up to the matching brace at depth 0
        depth, i = 1, start
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        body = src[start:i - 1]
        line0 = src[:fn.start()].count("\n") + 1
        nodes, shapes, cases = [], [], []
        latest = {}   # Rust lets a test shadow `let mul = ..`; give every definition a unique name

        def define(nm):
            k = f"{nm}#{len(nodes)}"
            latest[nm] = k
            return k

        def ref(nm):
            return latest[nm]
        cur_root = None
        ok = True
        last_case = None
        for st in split_statements(body):
            m = re.fullmatch(r"let (?:mut )?(\w+) = Context::new\(\);", st)
            if m:
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.([xyz])\(\);", st)
            if m:
                nodes.append([define(m.group(1)), "var", [m.group(2)]])
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.constant\((%s)\);" % NUM, st)
            if m:
                nodes.append([define(m.group(1)), "const", [num(m.group(2))]])
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.(\w+)\((.*)\)\.unwrap\(\);", st)
            if m:
                args = []
                for a in [x.strip() for x in m.group(3).split(",")]:
                    args.append(ref(a) if re.fullmatch(r"[A-Za-z_]\w*", a) else num(a))
                nodes.append([define(m.group(1)), m.group(2), args])
                continue
            m = re.fullmatch(r"let (?:shape|s) = F::new\(&ctx, &\[(\w+)\]\)\.unwrap\(\);", st)
            if m:
                cur_root = ref(m.group(1))
                continue
            if re.match(r"let (tape|vs|mut eval|eval) = ", st) or st.startswith("use "):
                continue
            m = re.search(r"eval\s*\.eval\(&tape, &(vs\((.*?)\)|\[(.*?)\])\s*\)", st)
            if m and cur_root:
                try:
                    if m.group(2) is not None:
                        ivs = [interval(t) for t in re.findall(r"\[[^\]]*\]", m.group(2))]
                    else:
                        ivs = [interval(t.replace(".into()", "")) for t in re.findall(r"\[[^\[\]]*\](?:\.into\(\))?", m.group(3))]
                except ValueError:
                    ok = False
                    break
                last_case = {"root": cur_root, "inputs": ivs}
                cases.append(last_case)
                e = re.search(r"\.0\[0\]\s*,\s*(\[[^\]]*\])\.into\(\)\s*\)", st)
                if e:
                    try:
                        last_case["expect"] = interval(e.group(1))
                    except ValueError:
                        cases.pop()   # expectation is an expression (e.g. `1.0_f32.sin()`): not a literal golden
                        last_case = None
                continue
            if last_case is not None:
                e = re.fullmatch(r"assert_eq!\(\s*\w+\[0\]\s*,\s*(\[[^\]]*\])\.into\(\)\s*\);", st)
                if e:
                    try:
                        last_case["expect"] = interval(e.group(1))
                    except ValueError:
                        pass
                    continue
                if re.fullmatch(r"assert!\(\w+(\[0\])?\.(lower|upper)\(\)\.is_nan\(\)\);", st):
                    last_case["expect"] = ["nan", "nan"]
                    continue
                if re.fullmatch(r"assert!\((data|trace)\.is_none\(\)\);", st):
                    last_case["trace"] = None
                    continue
                e = re.fullmatch(r"assert_eq!\(\s*(?:data|trace)\.unwrap\(\)\.as_ref\(\)\s*,\s*&\[(.*?)\]\s*,?\s*\);", st)
                if e:
                    last_case["trace"] = [c.strip().replace("Choice::", "") for c in e.group(1).split(",") if c.strip()]
                    continue
            ok = False
            break
        cases = [c for c in cases if "expect" in c]
        if ok and cases:
            tests[name] = {"source": f"{os.path.relpath(path, REF)}:{line0}", "nodes": nodes, "cases": cases}
        else:
            skipped += 1
    return tests, skipped


tests, skipped = parse_tests(os.path.join(REF, "fidget-core/src/eval/test/interval.rs"), "test_i_")
json.dump(tests, open(os.path.join(HERE, "interval_known_answers.json"), "w"), indent=1)
print("interval tests transcribed:", len(tests), "cases:", sum(len(t["cases"]) for t in tests.values()),
      "skipped fns:", skipped)
print(sorted(tests))


# ---------------------------------------------------------------------------
# grad_slice.rs known answers: `Self::eval_xyz(&tape, &[x], &[y], &[z])[0]` == `Grad::new(v, dx, dy, dz)`
def parse_grad_tests(path):
    src = open(path).read()
    tests, skipped = {}, 0
    for fn in re.finditer(r"pub fn (test_g_\w+)\(\) \{", src):
        name = fn.group(1)
        start = fn.end()
        depth, i = 1, start
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        body = src[start:i - 1]
        line0 = src[:fn.start()].count("\n") + 1
        nodes, cases, latest = [], [], {}
        cur_root, ok = None, True
        for st in split_statements(body):
            if re.fullmatch(r"let (?:mut )?\w+ = Context::new\(\);", st):
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.([xyz])\(\);", st)
            if m:
                k = f"{m.group(1)}#{len(nodes)}"; latest[m.group(1)] = k
                nodes.append([k, "var", [m.group(2)]])
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.constant\((%s)\);" % NUM, st)
            if m:
                k = f"{m.group(1)}#{len(nodes)}"; latest[m.group(1)] = k
                nodes.append([k, "const", [num(m.group(2))]])
                continue
            m = re.fullmatch(r"let (\w+) = ctx\.(\w+)\((.*)\)\.unwrap\(\);", st)
            if m:
                try:
                    args = [latest[a] if re.fullmatch(r"[A-Za-z_]\w*", a) else num(a)
                            for a in [x.strip() for x in m.group(3).split(",")]]
                except (KeyError, ValueError):
                    ok = False
                    break
                k = f"{m.group(1)}#{len(nodes)}"; latest[m.group(1)] = k
                nodes.append([k, m.group(2), args])
                continue
            m = re.fullmatch(r"let (?:shape|s) = F::new\(&ctx, &\[(\w+)\]\)\.unwrap\(\);", st)
            if m:
                cur_root = latest.get(m.group(1))
                continue
            if re.match(r"let (tape|mut eval|eval) = ", st):
                continue
            m = re.fullmatch(r"assert_eq!\(\s*Self::eval_xyz\(&tape, &\[(%s)\], &\[(%s)\], &\[(%s)\]\)\[0\]\s*,\s*"
                             r"Grad::new\((%s), (%s), (%s), (%s)\)\s*,?\s*\);" % ((NUM,) * 7), st)
            if m and cur_root:
                g = [num(v) for v in m.groups()]
                cases.append({"root": cur_root, "xyz": g[:3], "expect": g[3:]})
                continue
            ok = False
            break
        if ok and cases:
            tests[name] = {"source": f"{os.path.relpath(path, REF)}:{line0}", "nodes": nodes, "cases": cases}
        else:
            skipped += 1
    return tests, skipped


gtests, gskipped = parse_grad_tests(os.path.join(REF, "fidget-core/src/eval/test/grad_slice.rs"))
json.dump(gtests, open(os.path.join(HERE, "grad_known_answers.json"), "w"), indent=1)
print("grad tests transcribed:", len(gtests), "cases:", sum(len(t["cases"]) for t in gtests.values()), "skipped fns:", gskipped)
print(sorted(gtests))
