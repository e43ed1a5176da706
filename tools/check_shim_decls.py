#!/usr/bin/env python3
"""Token-for-token check of the drop-in headers against the reference's public declarations.

For every class / struct the shims replace, the PUBLIC member declarations (functions, constructors, data members; inline bodies
dropped) are tokenised and compared with the declaration of the same name and arity in the reference header:

    eorb_slam_b200/shim/ORBextractor.h            <->  include/ORBextractor.h             ORBxParams, ORBextractor
    eorb_slam_b200/shim/ORBmatcher_b200.h         <->  include/ORBmatcher.h               ORBmatcher (the members the shim declares)
    eorb_slam_b200/shim/EventConversion_b200.h    <->  include/Event/EventConversion.h    EvImConverter

The reference tree does not travel with the repository, so its side of the comparison is committed as a digest
(tests/golden/ref_decls.json: name, arity and sha1 of the normalised token string of every public declaration, no source text).
`--update` rewrites the digest from the reference tree (authoring container only); without the flag the shims are checked against
the committed digest, and, where the tree exists, the digest is re-derived and must be unchanged.

Exit status 0 = every shim declaration that exists in the reference matches it token for token, every reference declaration of a
replaced class is present in the shim or listed under ALLOW_MISSING, and every shim addition is listed under ALLOW_EXTRA.
"""
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EORB_REFERENCE", "/root/reference")
DIGEST = os.path.join(ROOT, "tests", "golden", "ref_decls.json")

PAIRS = [
    # (shim header, reference header, [class names])
    ("eorb_slam_b200/shim/ORBextractor.h", "include/ORBextractor.h", ["ORBxParams", "ORBextractor"]),
    # the matcher shims (ORBmatcher_b200.cc, ORBmatcher_guided_b200.cc) DEFINE members of the reference's own class against the reference's
    # own header; the declarations they are compiled against in this repository's mock build are the ones compared here
    ("eorb_slam_b200/shim/cv_mock/ref_mock.h", "include/ORBmatcher.h", ["ORBmatcher"]),
    ("eorb_slam_b200/shim/EventConversion_b200.h", "include/Event/EventConversion.h", ["EvImConverter"]),
]
# reference members the shims deliberately do not declare (kept by the reference's own translation unit)
ALLOW_MISSING = {
    "ORBextractor": set(),
    "ORBmatcher": None,          # None: the shim replaces a SUBSET of this class; only the members it declares are compared
    # the depth-map overload is dead code in the reference (no caller); the Jacobian takes its g2o vertex as plain doubles in the shim
    "EvImConverter": {"ev2mci_gg_f", "ev2mci_gg_f_jac"},
}
# members the shims add on top of the reference surface
ALLOW_EXTRA = {
    "ORBextractor": {"~ORBextractor/0", "ORBextractor/1:deleted-copy", "operator=/1", "mbDownloadPyramid/data", "SetDevice/1", "GetDevice/0"},
    "ORBxParams": set(), "ORBmatcher": None, "EvImConverter": None,
}


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def class_body(src, name):
    m = re.search(r"\b(class|struct)\s+%s\b[^;{]*\{" % re.escape(name), src)
    if not m:
        return None, None
    i = m.end()
    depth, j = 1, i
    while depth and j < len(src):
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        j += 1
    return m.group(1), src[i:j - 1]


def public_part(kind, body):
    """concatenation of the public sections of a class body"""
    out, cur = [], "public" if kind == "struct" else "private"
    pos = 0
    for m in re.finditer(r"\b(public|protected|private)\s*:", body):
        if cur == "public":
            out.append(body[pos:m.start()])
        cur, pos = m.group(1), m.end()
    if cur == "public":
        out.append(body[pos:])
    return "\n".join(out)


def drop_bodies(s):
    """inline function bodies and constructor initialiser lists -> ';'"""
    out, depth, i = [], 0, 0
    while i < len(s):
        c = s[i]
        if c == "{":
            if depth == 0:
                # drop a constructor initialiser list that precedes the body: the first ':' at parenthesis depth 0 of the
                # current statement that is not part of '::'
                txt = "".join(out)
                st = txt.rfind(";") + 1
                pd, cut = 0, -1
                for q in range(st, len(txt)):
                    ch = txt[q]
                    if ch in "(<":
                        pd += 1
                    elif ch in ")>":
                        pd -= 1
                    elif ch == ":" and pd == 0 and txt[q - 1:q] != ":" and txt[q + 1:q + 2] != ":":
                        cut = q
                        break
                if cut >= 0:
                    out = list(txt[:cut])
                out.append(";")
            depth += 1
        elif c == "}":
            depth -= 1
        elif depth == 0:
            out.append(c)
        i += 1
    return "".join(out)


TOKEN = re.compile(r"[A-Za-z_]\w*|::|->|<<|>>|\d+(?:\.\d*)?f?|[^\s\w]")


def declarations(src, cls):
    kind, body = class_body(src, cls)
    if body is None:
        return {}
    pub = drop_bodies(public_part(kind, body))
    decls = {}
    for stmt in pub.split(";"):
        toks = TOKEN.findall(stmt)
        if not toks or toks[0] in ("friend", "using", "typedef", "enum") or toks == ["}"]:
            continue
        if "(" in toks:
            k = toks.index("(")
            name = toks[k - 1]
            if name == "=" and k >= 2 and toks[k - 2] == "operator":
                name = "operator="
            elif k >= 2 and toks[k - 2] == "operator":
                name = "operator" + toks[k - 1]
            elif toks[k - 1] == ")" :      # operator()( ... )
                name = "operator()"
                k = k  # parameters start at this '('
            if k >= 2 and toks[k - 2] == "~":
                name = "~" + name
            # arity = top-level commas between the matching parentheses
            depth, n, empty = 0, 0, True
            for t in toks[k:]:
                if t in "(<[":
                    depth += 1
                elif t in ")>]":
                    depth -= 1
                    if depth == 0:
                        break
                elif depth == 1:
                    empty = False
                    if t == ",":
                        n += 1
            arity = 0 if empty else n + 1
            key = "%s/%d" % (name, arity)
        else:
            if "=" in toks:
                toks = toks[:toks.index("=")]          # a default member initialiser is not part of the declaration compared
            key = "%s/data" % toks[-1] if toks[-1] != "]" else "%s/data" % toks[toks.index("[") - 1]
        norm = " ".join(toks)
        # overloads with equal arity are told apart by a stable suffix
        base, i = key, 1
        while key in decls and decls[key] != norm:
            i += 1
            key = "%s#%d" % (base, i)
        decls[key] = norm
    return decls


def sha(s):
    return hashlib.sha1(s.encode()).hexdigest()


def reference_digest():
    d = {}
    for _, ref_hdr, classes in PAIRS:
        src = strip_comments(open(os.path.join(REF, ref_hdr)).read())
        for c in classes:
            d[c] = {k: sha(v) for k, v in sorted(declarations(src, c).items())}
    return d


def main():
    have_ref = os.path.exists(os.path.join(REF, "include", "ORBextractor.h"))
    if "--update" in sys.argv:
        assert have_ref, "needs the reference tree"
        json.dump(reference_digest(), open(DIGEST, "w"), indent=1, sort_keys=True)
        print("wrote", DIGEST)
    ref = json.load(open(DIGEST))
    problems = []
    if have_ref and reference_digest() != ref:
        problems.append("tests/golden/ref_decls.json is stale: re-run tools/check_shim_decls.py --update")
    for shim_hdr, _, classes in PAIRS:
        src = strip_comments(open(os.path.join(ROOT, shim_hdr)).read())
        for c in classes:
            mine = declarations(src, c)
            theirs = ref.get(c, {})
            if not mine:
                problems.append("%s: class %s not found" % (shim_hdr, c))
                continue
            matched = 0
            for key, norm in mine.items():
                name = key.split("#")[0]
                cands = [k for k in theirs if k.split("#")[0] == name]
                if not cands:
                    if ALLOW_EXTRA.get(c) is not None and name not in {e.split(":")[0] for e in ALLOW_EXTRA[c]}:
                        problems.append("%s::%s is not in the reference and not an allowed addition" % (c, key))
                    continue
                if sha(norm) in {theirs[k] for k in cands}:
                    matched += 1
                else:
                    allowed = {e.split(":")[0] for e in (ALLOW_EXTRA.get(c) or set())}
                    if name not in allowed:
                        problems.append("%s::%s differs from the reference declaration token for token:\n      %s" % (c, key, norm))
            if ALLOW_MISSING.get(c) is not None:
                mine_names = {k.split("#")[0] for k in mine}
                skip = {e.split(":")[0] for e in ALLOW_MISSING[c]}      # bare member names
                for key in theirs:
                    name = key.split("#")[0]
                    if name not in mine_names and name.split("/")[0] not in skip:
                        problems.append("%s::%s of the reference is missing from the shim" % (c, key))
                # overload counts
                for name in mine_names:
                    nr = len([k for k in theirs if k.split("#")[0] == name]); nm = len([k for k in mine if k.split("#")[0] == name])
                    if nr and nm < nr and name.split("/")[0] not in skip:
                        problems.append("%s::%s: the reference has %d overload(s), the shim %d" % (c, name, nr, nm))
            print("%-14s %2d shim declarations, %2d identical to the reference's, %2d in the reference class" % (c, len(mine), matched, len(theirs)))
    for p in problems:
        print("PROBLEM:", p)
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
