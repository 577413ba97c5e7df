"""ORACLE PINNING — test infrastructure only.

Mechanical C# -> C++ transpiler for the straight-line constraint / wide-math code of the reference (bepuphysics2). It reads the reference's
OWN source files where they lie (/root/reference, read-only), rewrites syntax only (parameter modifiers, `out var` declarations, static member
access, literals, generics -> templates) and writes one generated header into oracle/_ref/ (git-ignored: no reference source is committed).
No arithmetic is re-expressed by hand: every operator and every call in the generated C++ is the one the C# text has, in the same order; the
only hand-written code underneath is `ref_runtime.h` (System.Numerics.Vector<T> lane semantics: IEEE fp32 + - * / sqrt min max abs, compares
producing all-ones masks, bitwise select). The hand-written oracle (oracle/*.h) is then checked against this library bit for bit
(tests/test_oracle_pinned_to_reference.py), which ties the oracle's arithmetic to the C# text instead of to its author's reading of it.

Usage: python oracle/ref_transpile/cs2cpp.py <reference root> <output header> [--harness <output cpp>]
"""
import json
import os
import re
import sys

# ---------------------------------------------------------------------------------------------------------------------------------
# What to transpile: (file, [type names]) — None = every struct/class in the file that is not a description / type processor.
UTIL = "BepuUtilities/"
CON = "BepuPhysics/Constraints/"
SOURCES = [
    (UTIL + "MathHelper.cs", ["MathHelper"]),
    (UTIL + "Vector2Wide.cs", None), (UTIL + "Vector3Wide.cs", None), (UTIL + "Vector4Wide.cs", None), (UTIL + "QuaternionWide.cs", None),
    (UTIL + "Matrix2x3Wide.cs", None), (UTIL + "Matrix3x3Wide.cs", None), (UTIL + "Matrix2x2Wide.cs", None),
    (UTIL + "Symmetric2x2Wide.cs", None), (UTIL + "Symmetric3x3Wide.cs", None), (UTIL + "Symmetric4x4Wide.cs", None),
    (UTIL + "Symmetric5x5Wide.cs", None), (UTIL + "Symmetric6x6Wide.cs", None),
    ("BepuPhysics/BodyProperties.cs", ["BodyVelocityWide", "BodyInertiaWide"]),
    ("BepuPhysics/Helpers.cs", ["Helpers"]),
    ("BepuPhysics/PoseIntegrator.cs", ["PoseIntegration"]),
    (CON + "SpringSettings.cs", ["SpringSettingsWide"]), (CON + "ServoSettings.cs", ["ServoSettingsWide"]), (CON + "MotorSettings.cs", ["MotorSettingsWide"]),
    (CON + "InequalityHelpers.cs", None),
    ("BepuPhysics/CollisionDetection/PairMaterialProperties.cs", ["MaterialPropertiesWide"]),
    (CON + "Contact/PenetrationLimit.cs", None), (CON + "Contact/PenetrationLimitOneBody.cs", None),
    (CON + "Contact/TangentFriction.cs", None), (CON + "Contact/TangentFrictionOneBody.cs", None),
    (CON + "Contact/TwistFriction.cs", None), (CON + "Contact/TwistFrictionOneBody.cs", None),
    (CON + "Contact/ContactConvexCommon.cs", None),
    (CON + "Contact/ContactConvexTypes.cs", None), (CON + "Contact/ContactNonconvexCommon.cs", None), (CON + "Contact/ContactNonconvexTypes.cs", None),
]
JOINT_FILES = ["BallSocketShared", "BallSocket", "AngularHinge", "AngularSwivelHinge", "SwingLimit", "TwistServo", "TwistLimit", "TwistMotor", "AngularServo", "AngularMotor",
               "Weld", "VolumeConstraint", "DistanceServo", "DistanceLimit", "CenterDistanceConstraint", "AreaConstraint", "PointOnLineServo", "LinearAxisServo",
               "LinearAxisMotor", "LinearAxisLimit", "AngularAxisMotor", "OneBodyAngularServo", "OneBodyAngularMotor", "OneBodyLinearServo", "OneBodyLinearMotor",
               "SwivelHinge", "Hinge", "BallSocketMotor", "BallSocketServo", "AngularAxisGearMotor", "CenterDistanceLimit"]
SOURCES += [(CON + f + ".cs", None) for f in JOINT_FILES]
# PredictBoundingBoxes: the wide GetBounds of the convex primitives and the expansion helpers
SOURCES += [("BepuPhysics/BoundingBoxHelpers.cs", ["BoundingBoxHelpers"]), ("BepuPhysics/Collidables/Capsule.cs", ["CapsuleWide"]),
            ("BepuPhysics/Collidables/Box.cs", ["BoxWide"]), ("BepuPhysics/Collidables/Cylinder.cs", ["CylinderWide"])]

# Types of which only the named methods are taken (the rest of the type needs shapes, rays, narrow types ... outside this path).
ONLY_METHODS = {"CapsuleWide": {"GetBounds"}, "BoxWide": {"GetBounds"}, "CylinderWide": {"GetBounds"},
                "BoundingBoxHelpers": {"GetAngularBoundsExpansion", "GetBoundsExpansion"}}
# Types never emitted: descriptions (narrow <-> wide scatter code), type processors, interfaces.
SKIP_TYPE = re.compile(r"TypeProcessor$|^I[A-Z]\w*$|^NonconvexConstraintHelpers$|^ConvexConstraintHelpers$")
# A member is skipped (not an error) when its text needs something outside the hot-path arithmetic.
UNSUPPORTED = re.compile(
    r"GatherScatter|GetFirst\(|GetOffsetInstance|TypeBatch\b|Buffer<|BufferPool|IUnmanagedMemoryPool|\bstring\b|stackalloc|\bfixed\b|\bthrow\b|Span<|typeof|sizeof|nameof|"
    r"Vector128|Vector256|Vector512|\bAvx|\bSse|\bFma\b|\bAdvSimd|ArmBase|\bBitOperations|\bMathChecker|Validate|"
    r"ConstraintDescription|\bBodyHandle\b|\bSolver\b|\bBodies\b|\[\s*\d+\s*\]|\.ToString|\bConsole\b|\bList<|\bint\*|\bfloat\*|\bvoid\*|\bbyte\*|\bIEquatable|GetHashCode|\bobject\b")
NARROW = {"Vector2", "Vector3", "Vector4", "Quaternion", "Matrix3x3", "Matrix", "Matrix2x2", "Matrix2x3", "Symmetric3x3", "Symmetric2x2", "RigidPose", "BodyInertia", "BodyVelocity",
          "SpringSettings", "ServoSettings", "MotorSettings", "Matrix4x4", "Symmetric4x4", "Symmetric5x5", "Symmetric6x6", "PairMaterialProperties", "AffineTransform", "BoundingBox"}
MODS = r"(?:(?:public|private|internal|protected|static|unsafe|readonly|override|virtual|sealed|new|extern|partial|abstract)\s+)*"
CPP_TYPES = {"uint": "uint32_t", "ulong": "uint64_t", "long": "int64_t", "byte": "uint8_t", "sbyte": "int8_t", "ushort": "uint16_t", "short": "int16_t"}


def strip_comments(s):
    out, i, n = [], 0, len(s)
    while i < n:
        if s.startswith("//", i):
            j = s.find("\n", i)
            i = n if j < 0 else j
        elif s.startswith("/*", i):
            i = s.find("*/", i) + 2
        elif s[i] == '"':
            j = i + 1
            while s[j] != '"' or s[j - 1] == "\\":
                j += 1
            out.append('""')
            i = j + 1
        else:
            out.append(s[i])
            i += 1
    s = "".join(out)
    s = re.sub(r"^[ \t]*#.*$", "", s, flags=re.M)                      # preprocessor lines (#if DEBUG bodies are dropped below with Debug.Assert)
    s = re.sub(r"^[ \t]*\[[^\]\n]*\][ \t]*$", "", s, flags=re.M)       # attributes on their own line
    s = re.sub(r"\[MethodImpl\([^\]]*\)\]", "", s)
    return s


def match_brace(s, i, open_c="{", close_c="}"):
    depth = 0
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def split_top(s, sep=","):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "(<[{":
            depth += 1
        elif ch in ")>]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    return [p.strip() for p in parts]


class Method:
    def __init__(self):
        self.name = self.ret = self.body = ""
        self.static = False
        self.params = []  # (modifier, type, name)
        self.kind = "method"  # method | operator | ctor | property
        self.generic = None


class Type:
    def __init__(self, name, generic):
        self.name, self.generic = name, generic
        self.fields = []   # (type, name, static_const_init or None)
        self.methods = []
        self.props = set()


def find_types(src):
    pos, out = 0, []
    rx = re.compile(r"\b(struct|class|interface)\s+(\w+)\s*(<[^>{]*>)?\s*(:[^{]*?)?(\s*where[^{]*)?\{")
    while True:
        m = rx.search(src, pos)
        if not m:
            return out
        end = match_brace(src, m.end() - 1)
        out.append((m.group(1), m.group(2), m.group(3), src[m.end():end]))
        pos = end + 1


def split_members(body):
    members, i, n, start = [], 0, len(body), 0
    paren = 0
    seen_eq = False
    while i < n:
        c = body[i]
        if c == "(":
            paren += 1
        elif c == ")":
            paren -= 1
        elif c == "=" and paren == 0 and body[i + 1] not in "=>" and body[i - 1] not in "=!<>+-*/|&":
            seen_eq = True
        elif c == ";" and paren == 0:
            members.append(body[start:i + 1].strip())
            start, seen_eq = i + 1, False
        elif c == "{" and paren == 0:
            j = match_brace(body, i)
            if seen_eq:
                i = j
            else:
                k = j + 1
                while k < n and body[k].isspace():
                    k += 1
                if k < n and body[k] == "=":     # auto-property initializer
                    i = body.find(";", k)
                    members.append(body[start:i + 1].strip())
                else:
                    members.append(body[start:j + 1].strip())
                    i = j
                start, seen_eq = i + 1, False
        i += 1
    return [m for m in members if m]


def strip_intrinsic_branches(text):
    """`if (Avx.IsSupported && ...) {A} else if (Sse.IsSupported ...) {B} else {C}` -> `{C}`: hardware-intrinsic fast paths (bit-identical lane
    arithmetic by contract, except rcpps/rsqrtps) are removed and the portable branch the reference itself carries is what gets transpiled."""
    rx = re.compile(r"\bif\s*\(\s*(?:Avx2?|Sse\d*|Fma|AdvSimd|ArmBase|Vector(?:128|256|512))\.(?:X64\.)?Is(?:Supported|HardwareAccelerated)")
    while True:
        m = rx.search(text)
        if not m:
            return text
        close = match_brace(text, text.index("(", m.start()), "(", ")")
        k = close + 1
        while text[k].isspace():
            k += 1
        end = match_brace(text, k) if text[k] == "{" else text.index(";", k)
        rest = text[end + 1:]
        rest = re.sub(r"^\s*else\b", " ", rest, count=1)
        text = text[:m.start()] + rest


def parse_params(text):
    params = []
    for p in split_top(text):
        p = re.sub(r"\s*=\s*[^,]+$", "", p)  # default value
        m = re.match(r"^(?:this\s+)?(?:(in|ref|out|params)\s+)?(.+?)\s+(\w+)$", p, flags=re.S)
        if not m:
            return None
        params.append((m.group(1) or "", re.sub(r"\s+", "", m.group(2)), m.group(3)))
    return params


def parse_type(kind, name, generic, body):
    t = Type(name, [g.strip() for g in generic.strip("<>").split(",")] if generic else None)
    for mem in split_members(body):
        if re.match(r"^" + MODS + r"(struct|class|interface|enum|delegate)\b", mem):
            continue
        mem = strip_intrinsic_branches(mem)
        # text before the body
        brace = None
        depth = 0
        for idx, ch in enumerate(mem):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif depth == 0 and (ch == "{" or mem.startswith("=>", idx)):
                brace = idx
                break
        header = mem[:brace].strip() if brace is not None else mem.rstrip(";").strip()
        header1 = re.sub(r"\s+", " ", header)
        if UNSUPPORTED.search(mem):
            continue
        words = set(re.findall(r"(?<![\w.])[A-Z]\w*\b(?!\s*[=;,)])", mem))  # identifiers in type / static-call position, not field names
        if words & NARROW:
            continue
        if "(" in header1 and brace is not None or ("(" in header1 and "=" not in header1.split("(")[0]):
            meth = Method()
            mc = re.match(r"^(" + MODS + r")" + re.escape(name) + r"\s*\((.*)\)\s*$", header1)
            mo = re.match(r"^(" + MODS + r")(?:(implicit|explicit)\s+)?(.+?)\s*\boperator\s*(\S+?)\s*\((.*)\)$", header1)
            mm = re.match(r"^(" + MODS + r")(.+?)\s+(\w+)\s*(<[^()]*>)?\s*\((.*)\)\s*(where .*)?$", header1)
            if mc:
                meth.kind, meth.name, ptext = "ctor", name, mc.group(2)
            elif mo:
                if mo.group(2):
                    continue
                meth.kind, meth.ret, meth.name, ptext, meth.static = "operator", mo.group(3).strip(), "operator" + mo.group(4), mo.group(5), True
            elif mm:
                meth.ret, meth.name, ptext = mm.group(2).strip(), mm.group(3), mm.group(5)
                meth.static = bool(re.search(r"\bstatic\b", mm.group(1)))
                if mm.group(4):
                    meth.generic = [g.strip() for g in mm.group(4).strip("<>").split(",")]
            else:
                continue
            if name in ONLY_METHODS and meth.name not in ONLY_METHODS[name]:
                continue
            meth.params = parse_params(ptext)
            if meth.params is None or brace is None:
                continue
            rest = mem[brace:]
            meth.body = ("{ return " + rest[2:].strip().rstrip(";") + "; }") if rest.startswith("=>") else rest
            if meth.body.strip() in ("{ return ; }",):
                continue
            t.methods.append(meth)
        elif brace is not None:
            # property: only static/instance expression-bodied getters are kept
            mp = re.match(r"^(" + MODS + r")(.+?)\s+(\w+)$", header1)
            rest = mem[brace:]
            if mp and rest.startswith("=>"):
                meth = Method()
                meth.kind, meth.ret, meth.name = "property", mp.group(2).strip(), mp.group(3)
                meth.static = bool(re.search(r"\bstatic\b", mp.group(1)))
                meth.body = "{ return " + rest[2:].strip().rstrip(";") + "; }"
                t.methods.append(meth)
                t.props.add(meth.name)
        else:
            mf = re.match(r"^(" + MODS + r")(const\s+)?([\w<>,\.\[\]]+)\s+(.+)$", header1)
            if not mf:
                continue
            is_const = bool(mf.group(2)) or ("static" in mf.group(1) and "readonly" in mf.group(1))
            ftype = mf.group(3)
            for decl in split_top(mf.group(4)):
                dm = re.match(r"^(\w+)(?:\s*=\s*(.+))?$", decl, flags=re.S)
                if dm:
                    t.fields.append((ftype, dm.group(1), dm.group(2) if is_const else None, is_const))
    return t


# ---------------------------------------------------------------------------------------------------------------------------------
class Transpiler:
    def __init__(self, root):
        self.root = root
        self.types = {}
        self.order = []
        self.processors = {}  # XTypeProcessor -> [prestep struct, impulse struct, functions struct, access filters...]

    def load(self):
        for rel, wanted in SOURCES:
            path = os.path.join(self.root, rel)
            if not os.path.exists(path):
                continue
            src = strip_comments(open(path, encoding="utf-8-sig").read())
            for pm in re.finditer(r"class\s+(\w+TypeProcessor)\s*:\s*\w+\s*<", src):
                close = match_brace(src, pm.end() - 1, "<", ">")
                self.processors[pm.group(1)] = split_top(src[pm.end():close])
            static_imports = [x.split(".")[-1] for x in re.findall(r"^\s*using\s+static\s+([\w.]+)\s*;", src, flags=re.M)]
            found = []
            for kind, name, generic, body in find_types(src):
                # nested structs are hoisted to namespace scope as Owner__Nested (C# resolves the unqualified name inside the owner)
                for nk, nn, ng, nbody in find_types(body):
                    if nk == "struct" and not SKIP_TYPE.search(nn) and kind != "interface":
                        mangled = name + "__" + nn
                        body = re.sub(r"\b(?:public\s+|private\s+|internal\s+)?struct\s+%s\b" % nn, "struct " + mangled, body)
                        body = re.sub(r"(?<![\w.])%s\b" % nn, mangled, body)
                        found.append(("struct", mangled, ng, find_types(body)[[x[1] for x in find_types(body)].index(mangled)][3]))
                found.append((kind, name, generic, body))
            for kind, name, generic, body in found:
                if kind == "interface" or SKIP_TYPE.search(name):
                    continue
                if wanted is not None and name not in wanted:
                    continue
                if wanted is None and name in NARROW:
                    continue
                if wanted is None and "__" not in name and not re.search(r"Wide$|Functions$|PrestepData$|AccumulatedImpulses$|^PenetrationLimit|^TangentFriction|^TwistFriction|Helpers$|^FrictionHelpers$|Shared$|^NonconvexContactPrestepData$|^MaterialPropertiesWide$", name):
                    continue
                t = parse_type(kind, name, generic, body)
                t.source = rel
                t.static_imports = static_imports
                if name in self.types:
                    continue
                self.types[name] = t
                self.order.append(name)

    # -- type names ---------------------------------------------------------------------------------------------------------------
    def cpp_type(self, t):
        t = t.strip()
        m = re.match(r"^ref\s+(.+)$", t)
        if m:
            return self.cpp_type(m.group(1)) + "&"
        return CPP_TYPES.get(t, t)

    def sig(self, owner, m, qualify, inline=False):
        ps = []
        for mod, ty, name in m.params:
            ty = self.cpp_type(ty)
            ps.append(("const %s& %s" if mod == "in" else "%s& %s" if mod in ("ref", "out") else "%s %s") % (ty, name))
        tmpl = ""
        if m.generic:
            tmpl = "template <%s> " % ", ".join("class " + g for g in m.generic)
        if inline:
            tmpl += "inline "
        scope = ""
        if qualify:
            scope = owner.name + ("<%s>" % ", ".join(owner.generic) if owner.generic else "") + "::"
        if m.kind == "ctor":
            return "%s%s%s(%s)" % (tmpl, scope, owner.name, ", ".join(ps))
        name = m.name
        static = "static " if (m.static and not qualify and m.kind != "operator") else ""
        const = ""
        return "%s%s%s %s%s(%s)%s" % (tmpl, static, self.cpp_type(m.ret), scope if m.kind != "operator" else "", name, ", ".join(ps), const)

    # -- `out var` type inference -------------------------------------------------------------------------------------------------
    def field_type(self, tname, field):
        t = self.types.get(tname)
        if not t:
            return None
        for ft, fn, _, _ in t.fields:
            if fn == field:
                return ft
        return None

    def expr_type(self, expr, env):
        expr = expr.strip()
        expr = re.sub(r"^(ref|in|out)\s+", "", expr)
        parts = expr.split(".")
        if not all(re.match(r"^\w+$", p) for p in parts):
            m = re.match(r"^new\s+([\w<>]+)\s*\(", expr)
            if m:
                return m.group(1)
            # arithmetic over resolvable operands: the widest operand type (C# operator overloads on the wide structs return the struct)
            leaves = [x for x in re.split(r"[-+*/()\s]+", expr) if x]
            if leaves and all(re.match(r"^[\w.]+$", x) for x in leaves) and len(leaves) > 1:
                tys = [self.expr_type(x, env) for x in leaves if not re.match(r"^[\d.]+f?$", x)]
                if tys and all(t is not None for t in tys):
                    wide = [t for t in tys if t not in ("Vector<float>", "float", "int")]
                    if not wide:
                        return "Vector<float>" if "Vector<float>" in tys else "float"
                    if len(set(wide)) == 1:
                        return wide[0]
            return None
        ty = env.get(parts[0])
        for p in parts[1:]:
            if ty is None:
                return None
            ty = self.field_type(ty, p)
        return ty

    def resolve_out_type(self, owner, callee, args, index, env):
        parts = callee.split(".")
        mname = parts[-1]
        cands = []
        search = []
        if len(parts) >= 2 and parts[-2] in self.types:
            search = [self.types[parts[-2]]]
        elif len(parts) == 1:
            search = [owner]
        else:
            oty = self.expr_type(".".join(parts[:-1]), env)
            search = [self.types[oty]] if oty in self.types else list(self.types.values())
        for t in search:
            for m in t.methods:
                if m.name == mname and len(m.params) == len(args):
                    cands.append(m)
        def arg_mod(a):
            mm = re.match(r"^(ref|out|in)\s", a.strip())
            return mm.group(1) if mm else ""
        c2 = [m for m in cands if all((arg_mod(a) == p[0]) or (arg_mod(a) == "" and p[0] in ("", "in")) for a, p in zip(args, m.params))]
        cands = c2 or cands
        types = {m.params[index][1] for m in cands}
        if len(types) > 1:
            c3 = []
            for m in cands:
                ok = True
                for a, p in zip(args, m.params):
                    at = self.expr_type(a, env)
                    if at is not None and at != p[1]:
                        ok = False
                if ok:
                    c3.append(m)
            types = {m.params[index][1] for m in c3} or types
        if len(types) == 1:
            return types.pop()
        return None

    # -- body ---------------------------------------------------------------------------------------------------------------------
    def rewrite_calls(self, body, name_rx, fn):
        """Rewrites every call `name(args)` (balanced) through fn(match, [args]) -> replacement."""
        out, pos = [], 0
        rx = re.compile(name_rx + r"\s*\(")
        while True:
            m = rx.search(body, pos)
            if not m:
                out.append(body[pos:])
                return "".join(out)
            close = match_brace(body, m.end() - 1, "(", ")")
            args = split_top(body[m.end():close])
            out.append(body[pos:m.start()])
            out.append(fn(m, args))
            pos = close + 1

    def rewrite_calls_once(self, body, name_rx, fn):
        """Like rewrite_calls, but arguments are rewritten recursively first (nested calls) and matches inside the replacement are not revisited."""
        out, pos = [], 0
        rx = re.compile(name_rx + r"\s*\(")
        while True:
            m = rx.search(body, pos)
            if not m:
                out.append(body[pos:])
                return "".join(out)
            close = match_brace(body, m.end() - 1, "(", ")")
            inner = self.rewrite_calls_once(body[m.end():close], name_rx, fn)
            out.append(body[pos:m.start()])
            out.append(fn(m, split_top(inner)))
            pos = close + 1

    def translate_body(self, owner, m):
        body = m.body
        env = {p[2]: p[1] for p in m.params}
        for ft, fn, _, _ in owner.fields:
            env.setdefault(fn, ft)
        body = self.rewrite_calls(body, r"\bDebug\.Assert", lambda mm, a: "(void)0")
        if getattr(owner, "static_imports", None):
            own = {x.name for x in owner.methods}
            def qualify(mm, args):
                name = mm.group(1)
                text = mm.group(0) + ", ".join(args) + ")"
                if name in own or mm.start() > 0 and False:
                    return text
                cands = [(t, x) for t in owner.static_imports if t in self.types for x in self.types[t].methods if x.name == name and len(x.params) == len(args)]
                if not cands:
                    return text
                def fits(x):
                    for a, p in zip(args, x.params):
                        at = self.expr_type(a, env)
                        if at is not None and at != p[1]:
                            return False
                    return True
                good = [c for c in cands if fits(c[1])] or cands
                owners = {c[0] for c in good}
                if len(owners) != 1:
                    raise ValueError("ambiguous `using static` call %s in %s.%s" % (name, owner.name, m.name))
                return owners.pop() + "." + text
            prev = None
            while prev != body:
                prev = body
                body = self.rewrite_calls_once(body, r"(?<![\w.:])([A-Z]\w*)", qualify)
        body = re.sub(r"Unsafe\.SkipInit\(\s*out\s+([\w<>]+)\s+(\w+)\s*\)\s*;", r"\1 \2;", body)
        body = re.sub(r"Unsafe\.SkipInit\(\s*out\s+(\w[\w.]*)\s*\)\s*;", r"", body)
        for _ in range(4):
            body = self.rewrite_calls(body, r"\bUnsafe\.Add", lambda mm, a: "(&(%s))[%s]" % (re.sub(r"^ref\s+", "", a[0]), a[1]))
            body = self.rewrite_calls(body, r"\bUnsafe\.As<\s*[\w<>]+\s*,\s*([\w<>]+)\s*>", lambda mm, a: "(*reinterpret_cast<%s*>(&(%s)))" % (mm.group(1), re.sub(r"^ref\s+", "", a[0])))
        # explicit local declarations feed the environment
        for dm in re.finditer(r"(?:^|[;{}])\s*(?:ref\s+)?([A-Z]\w*(?:<\w+>)?)\s+(\w+)\s*(?:;|=)", body):
            env.setdefault(dm.group(2), dm.group(1))
        for dm in re.finditer(r"\bvar\s+(\w+)\s*=\s*new\s+([\w<>]+)\s*\(", body):
            env.setdefault(dm.group(1), dm.group(2))
        self.pending_vars = [(dm.start(), dm.group(1), dm.group(2)) for dm in re.finditer(r"\bvar\s+(\w+)\s*=\s*([^;]+);", body)]
        counter = [0]
        def discard(mm):
            counter[0] += 1
            return "out var discard_%d" % counter[0]
        body = re.sub(r"\bout\s+_(?=\s*[,)])", discard, body)
        # hoist `out var x` / `out T x` declarations in front of their statement
        while True:
            mo = re.search(r"\bout\s+(var|[A-Z]\w*(?:<\w+>)?|float|int|bool)\s+(\w+)\s*(?=[,)])", body)
            if not mo:
                break
            ty, name = mo.group(1), mo.group(2)
            for _, vn, vexpr in self.pending_vars:  # `var x = <arithmetic>` locals seen so far
                if vn not in env:
                    vt = self.expr_type(vexpr, env)
                    if vt:
                        env[vn] = vt
            if ty == "var":
                # enclosing call
                depth, i = 0, mo.start()
                while i >= 0:
                    if body[i] == ")":
                        depth += 1
                    elif body[i] == "(":
                        if depth == 0:
                            break
                        depth -= 1
                    i -= 1
                close = match_brace(body, i, "(", ")")
                args = split_top(body[i + 1:close])
                cm = re.search(r"([\w.<>]+)\s*$", body[:i])
                callee = re.sub(r"<[^>]*>$", "", cm.group(1)) if cm else ""
                index = next(k for k, a in enumerate(args) if re.match(r"^out\s+var\s+" + name + r"$", a.strip()))
                ty = self.resolve_out_type(owner, callee, args, index, env)
                if ty is None:
                    raise ValueError("cannot infer the type of `out var %s` in %s.%s (call %s)" % (name, owner.name, m.name, callee))
            env[name] = ty
            # statement start: after the previous ; { } at paren depth 0
            j, depth = mo.start(), 0
            while j > 0:
                ch = body[j - 1]
                if ch == ")":
                    depth += 1
                elif ch == "(":
                    depth -= 1
                elif ch in ";{}" and depth <= 0:
                    break
                j -= 1
            body = body[:j] + " %s %s;" % (self.cpp_type(ty), name) + body[j:mo.start()] + re.sub(r"\bout\s+(?:var|[\w<>]+)\s+" + name + r"\b", name, body[mo.start():], count=1)
        body = re.sub(r"\bref\s+var\s+(\w+)\s*=\s*ref\s+", r"auto& \1 = ", body)
        body = re.sub(r"\bref\s+([A-Z][\w<>]*)\s+(\w+)\s*=\s*ref\s+", r"\1& \2 = ", body)
        body = re.sub(r"\breturn\s+ref\s+", "return ", body)
        body = re.sub(r"(?<=[(,])\s*(?:ref|out|in)\s+(?=[\w(])", " ", body)
        body = re.sub(r"\bvar\s+(?=\w+\s*=)", "auto ", body)
        body = re.sub(r"\bnew\s+(?=[\w<>]+\s*\()", "", body)
        body = re.sub(r"\bdefault\s*\(\s*([\w<>]+)\s*\)", r"\1{}", body)
        body = re.sub(r"=\s*default\s*;", "= {};", body)
        return self.translate_expr(owner, body)

    def translate_expr(self, owner, body):
        names = set(self.types) | set(owner.generic or []) | {"MathF", "Math"}
        body = re.sub(r"(?<![\w.])Vector<(float|int)>\.", r"Vector<\1>::", body)
        body = re.sub(r"Vector<(float|int)>::(Zero|One|AllBitsSet)\b(?!\()", r"Vector<\1>::\2()", body)
        body = re.sub(r"(?<![\w.])Vector\.(?=\w+\s*[(<])", "VectorOps::", body)
        def static_access(mm):
            return mm.group(1) + "::" if mm.group(1) in names else mm.group(0)
        body = re.sub(r"(?<![\w.>])([A-Z]\w*)\.(?=\w)", static_access, body)
        # static properties of transpiled types
        for t in self.types.values():
            for p in t.props:
                body = re.sub(r"\b(\w+)::%s\b(?!\s*\()" % p, lambda mm: mm.group(0) + "()", body)
        body = re.sub(r"\bfloat\.PositiveInfinity\b", "INFINITY", body)
        body = re.sub(r"\bfloat\.MaxValue\b", "FLT_MAX", body)
        body = re.sub(r"\bfloat\.MinValue\b", "(-FLT_MAX)", body)
        body = re.sub(r"\bfloat\.Epsilon\b", "1.401298E-45f", body)
        body = re.sub(r"\bint\.MaxValue\b", "INT_MAX", body)
        body = re.sub(r"(?<![\w.])(?<![eE][-+])(\d+)[fF]\b", r"\1.0f", body)
        body = re.sub(r"(?<![\w.])(\d+\.\d+)[dD]\b", r"\1", body)
        body = re.sub(r"\bthis\.", "this->", body)
        body = re.sub(r"\bthis\b(?!->)", "(*this)", body)
        for k, v in CPP_TYPES.items():
            body = re.sub(r"\b%s\b" % k, v, body)
        return body

    # -- emission -----------------------------------------------------------------------------------------------------------------
    def emit(self):
        out = ["// GENERATED by oracle/ref_transpile/cs2cpp.py from the reference's C# sources. Do not edit, do not commit.", "#pragma once", '#include "ref_runtime.h"', "namespace bepu_ref {"]
        # order structs by field dependencies
        done, ordered = set(), []
        def visit(n, stack=()):
            if n in done or n not in self.types or n in stack:
                return
            for ft, _, _, const in self.types[n].fields:
                for dep in re.findall(r"\w+", ft):
                    visit(dep, stack + (n,))
            done.add(n)
            ordered.append(n)
        for n in self.order:
            visit(n)
        for n in ordered:
            t = self.types[n]
            out.append(("template <%s> " % ", ".join("class " + g for g in t.generic) if t.generic else "") + "struct %s;" % n)
        skipped = []
        for n in ordered:
            t = self.types[n]
            out.append(("template <%s> " % ", ".join("class " + g for g in t.generic) if t.generic else "") + "struct %s {" % n)
            for ft, fn, init, const in t.fields:
                if const:
                    out.append("    static constexpr %s %s = %s;" % (self.cpp_type(ft), fn, self.translate_expr(t, init)))
                else:
                    out.append("    %s %s;" % (self.cpp_type(ft), fn))
            if any(m.kind == "ctor" for m in t.methods):
                out.append("    %s() = default;" % n)
            for m in t.methods:
                if m.kind == "operator":
                    continue
                out.append("    " + (("static " if m.static else "") + "%s %s()" % (self.cpp_type(m.ret), m.name) if m.kind == "property" else self.sig(t, m, False)) + ";")
            out.append("};")
        for n in ordered:  # operator prototypes
            t = self.types[n]
            for m in t.methods:
                if m.kind == "operator":
                    out.append(self.sig(t, m, False, inline=True) + ";")
        for n in ordered:
            t = self.types[n]
            tmpl = "template <%s> " % ", ".join("class " + g for g in t.generic) if t.generic else ""
            scope = n + ("<%s>" % ", ".join(t.generic) if t.generic else "") + "::"
            for m in t.methods:
                try:
                    body = self.translate_body(t, m)
                except ValueError as e:
                    skipped.append(str(e))
                    body = None
                if body is None:
                    body = "{ static_assert(sizeof(%s) == 0, \"%s.%s could not be transpiled\"); }" % (n, n, m.name) if False else "{ REF_UNTRANSPILED(\"%s.%s\"); }" % (n, m.name)
                if m.kind == "operator":
                    for other in {x.name for x in t.methods if x.kind == "method" and x.static}:
                        body = re.sub(r"(?<![\w.:])%s\s*\(" % other, "%s::%s(" % (n, other), body)
                if m.kind == "property":
                    out.append("%sinline %s %s%s() %s" % (tmpl, self.cpp_type(m.ret), scope, m.name, body))
                elif m.kind == "operator":
                    out.append(self.sig(t, m, False, inline=True) + " " + body)
                else:
                    out.append(tmpl + self.sig(t, m, True, inline=True) + " " + body)
        out.append("}  // namespace bepu_ref")
        return "\n".join(out) + "\n", skipped


def harness(tr, layouts):
    """One lane-level entry point over every constraint type: same signature as the hand-written oracle's oracle_eval_lane."""
    out = ['// GENERATED by oracle/ref_transpile/cs2cpp.py. Do not edit, do not commit.', '#include "bepu_ref_generated.h"', "#include <cstring>", "using namespace bepu_ref;",
           "template <class T> static void load_rows(T& dst, const float* rows, int stride) { float* f = reinterpret_cast<float*>(&dst); for (size_t i = 0; i < sizeof(T) / sizeof(float); ++i) f[i] = rows[i * stride]; }",
           "template <class T> static void store_rows(const T& src, float* rows, int stride) { const float* f = reinterpret_cast<const float*>(&src); for (size_t i = 0; i < sizeof(T) / sizeof(float); ++i) rows[i * stride] = f[i]; }",
           "struct BodyIn { Vector3Wide pos; QuaternionWide q; BodyInertiaWide inertia; BodyVelocityWide v; };",
           "static BodyIn load_body(const float* f, const float* w) { BodyIn b; b.pos.X = f[0]; b.pos.Y = f[1]; b.pos.Z = f[2]; b.q.X = f[3]; b.q.Y = f[4]; b.q.Z = f[5]; b.q.W = f[6];",
           "    b.inertia.InverseInertiaTensor.XX = f[7]; b.inertia.InverseInertiaTensor.YX = f[8]; b.inertia.InverseInertiaTensor.YY = f[9]; b.inertia.InverseInertiaTensor.ZX = f[10];",
           "    b.inertia.InverseInertiaTensor.ZY = f[11]; b.inertia.InverseInertiaTensor.ZZ = f[12]; b.inertia.InverseMass = f[13];",
           "    b.v.Linear.X = w[0]; b.v.Linear.Y = w[1]; b.v.Linear.Z = w[2]; b.v.Angular.X = w[3]; b.v.Angular.Y = w[4]; b.v.Angular.Z = w[5]; return b; }",
           "static void store_velocity(const BodyIn& b, float* w) { w[0] = b.v.Linear.X.v; w[1] = b.v.Linear.Y.v; w[2] = b.v.Linear.Z.v; w[3] = b.v.Angular.X.v; w[4] = b.v.Angular.Y.v; w[5] = b.v.Angular.Z.v; }",
           'extern "C" int ref_eval_lane(int type_id, int stage, const float* body_states, float dt, float* prestep, float* impulses, float* velocities, int row_stride) {',
           "    BodyIn b[4];", "    switch (type_id) {"]
    covered = []
    for tid_s, tinfo in sorted(layouts["types"].items(), key=lambda kv: int(kv[0])):
        tid, nb = int(tid_s), tinfo["bodies"]
        decl = tr.processors.get(tinfo["processor"])
        if not decl:
            continue
        pre, acc, fn = decl[0], decl[1], re.sub(r"\s+", "", decl[2])
        assert pre == tinfo["prestep_struct"] and acc == tinfo["impulse_struct"], (pre, acc, tinfo)
        base = fn.split("<")[0]
        if base not in tr.types or pre not in tr.types or (acc not in tr.types and acc != "Vector<float>"):
            continue
        ft = tr.types[base]
        names = {mm.name for mm in ft.methods}
        if not {"WarmStart", "Solve"} <= names:
            continue
        covered.append(tid)
        bodies_ws = ", ".join("b[%d].pos, b[%d].q, b[%d].inertia" % (i, i, i) for i in range(nb))
        vels = ", ".join("b[%d].v" % i for i in range(nb))
        out.append("        case %d: {" % tid)
        out.append("            static_assert(sizeof(%s) == %d * sizeof(float) && sizeof(%s) == %d * sizeof(float), \"layout\");" % (pre, len(tinfo["prestep_rows"]), acc, len(tinfo["impulse_rows"])))
        out.append("            %s p; %s a; load_rows(p, prestep, row_stride); load_rows(a, impulses, row_stride);" % (pre, acc))
        out.append("            for (int s = 0; s < %d; ++s) b[s] = load_body(body_states + 14 * s, velocities + 6 * s);" % nb)
        out.append("            if (stage == 0) %s::WarmStart(%s, p, a, %s);" % (fn, bodies_ws, vels))
        out.append("            else if (stage == 1) %s::Solve(%s, dt, 1.0f / dt, p, a, %s);" % (fn, bodies_ws, vels))
        if "IncrementallyUpdateForSubstep" in names:
            out.append("            else %s::IncrementallyUpdateForSubstep(Vector<float>(dt), %s, p);" % (fn, vels))
        out.append("            store_rows(p, prestep, row_stride); store_rows(a, impulses, row_stride);")
        out.append("            for (int s = 0; s < %d; ++s) store_velocity(b[s], velocities + 6 * s);" % nb)
        out.append("            return 0; }")
    out += ["        default: return -1;", "    }", "}",
            "// PoseIntegration (BepuPhysics/PoseIntegrator.cs:L146-253), one call per function; operand layout of oracle_eval_integration.",
            "static QuaternionWide quat(const float* f) { QuaternionWide q; q.X = f[0]; q.Y = f[1]; q.Z = f[2]; q.W = f[3]; return q; }",
            "static Vector3Wide vec3(const float* f) { Vector3Wide v; v.X = f[0]; v.Y = f[1]; v.Z = f[2]; return v; }",
            "static Symmetric3x3Wide sym3(const float* f) { Symmetric3x3Wide m; m.XX = f[0]; m.YX = f[1]; m.YY = f[2]; m.ZX = f[3]; m.ZY = f[4]; m.ZZ = f[5]; return m; }",
            'extern "C" int ref_eval_integration(int op, const float* in, float* out) {',
            "    if (op == 0) { QuaternionWide r; PoseIntegration::Integrate(quat(in), vec3(in + 4), Vector<float>(in[7]), r); out[0] = r.X.v; out[1] = r.Y.v; out[2] = r.Z.v; out[3] = r.W.v; return 0; }",
            "    if (op == 1) { Symmetric3x3Wide r; PoseIntegration::RotateInverseInertia(sym3(in), quat(in + 6), r); out[0] = r.XX.v; out[1] = r.YX.v; out[2] = r.YY.v; out[3] = r.ZX.v; out[4] = r.ZY.v; out[5] = r.ZZ.v; return 0; }",
            "    if (op == 2) { Vector3Wide w = vec3(in + 16); PoseIntegration::IntegrateAngularVelocityConserveMomentum(quat(in), sym3(in + 4), sym3(in + 10), w); out[0] = w.X.v; out[1] = w.Y.v; out[2] = w.Z.v; return 0; }",
            "    if (op == 3) { Vector3Wide w = vec3(in + 10); PoseIntegration::IntegrateAngularVelocityConserveMomentumWithGyroscopicTorque(quat(in), sym3(in + 4), w, Vector<float>(in[13])); out[0] = w.X.v; out[1] = w.Y.v; out[2] = w.Z.v; return 0; }",
            "    return -1;", "}",
            "// MathHelper.Sin / Cos / Acos (BepuUtilities/MathHelper.cs:L274-362): the custom approximations every orientation integration goes through.",
            'extern "C" float ref_math(int fn, float x) { Vector<float> v(x); return fn == 0 ? MathHelper::Sin(v).v : fn == 1 ? MathHelper::Cos(v).v : MathHelper::Acos(v).v; }',
            "// Convex-primitive bounds of PredictBoundingBoxes: transpiled CapsuleWide / BoxWide / CylinderWide.GetBounds and BoundingBoxHelpers.GetAngularBoundsExpansion /",
            "// GetBoundsExpansion, glued exactly like BoundingBoxBatcher.ExecuteConvexBatch (Collidables/BoundingBoxBatcher.cs:L176-197) glues them. SphereWide.GetBounds",
            "// (Sphere.cs:L149-160: max = radius, min = -radius, no angular expansion) is three assignments and is written out here (its `new Vector3Wide(ref x)` overload pair has no C++ counterpart).",
            "// in: type, dims[3], margins {min, max}, allow, orientation[4], position[3], linear[3], angular[3] (velocity AFTER the callback), dt; out: min.xyz, margin, max.xyz.",
            'extern "C" int ref_convex_bounds(int type, const float* dims, const float* margins, int allow, const float* q, const float* pos, const float* lin, const float* ang, float dt, float* out) {',
            "    QuaternionWide orientations = quat(q); Vector3Wide positions = vec3(pos); BodyVelocityWide velocities; velocities.Linear = vec3(lin); velocities.Angular = vec3(ang);",
            "    Vector<float> maximumRadius, maximumAngularExpansion; Vector3Wide bundleMin, bundleMax; Vector<float> dtWide(dt);",
            "    if (type == 0) { maximumRadius = Vector<float>(0.0f); maximumAngularExpansion = Vector<float>(0.0f); Vector<float> Radius(dims[0]); Vector<float> negatedRadius = -Radius;",
            "        bundleMax.X = Radius; bundleMax.Y = Radius; bundleMax.Z = Radius; bundleMin.X = negatedRadius; bundleMin.Y = negatedRadius; bundleMin.Z = negatedRadius; }",
            "    else if (type == 1) { CapsuleWide s; s.Radius = dims[0]; s.HalfLength = dims[1]; s.GetBounds(orientations, 1, maximumRadius, maximumAngularExpansion, bundleMin, bundleMax); }",
            "    else if (type == 2) { BoxWide s; s.HalfWidth = dims[0]; s.HalfHeight = dims[1]; s.HalfLength = dims[2]; s.GetBounds(orientations, 1, maximumRadius, maximumAngularExpansion, bundleMin, bundleMax); }",
            "    else if (type == 4) { CylinderWide s; s.Radius = dims[0]; s.HalfLength = dims[1]; s.GetBounds(orientations, 1, maximumRadius, maximumAngularExpansion, bundleMin, bundleMax); }",
            "    else return -1;",
            "    Vector<float> angularSpeed; Vector3Wide::Length(velocities.Angular, angularSpeed); Vector<float> linearSpeed; Vector3Wide::Length(velocities.Linear, linearSpeed);",
            "    auto angularBoundsExpansion = BoundingBoxHelpers::GetAngularBoundsExpansion(angularSpeed, dtWide, maximumRadius, maximumAngularExpansion);",
            "    auto speculativeMargin = linearSpeed * dtWide + angularBoundsExpansion;",
            "    speculativeMargin = VectorOps::Max(Vector<float>(margins[0]), VectorOps::Min(Vector<float>(margins[1]), speculativeMargin));",
            "    auto maximumBoundsExpansion = VectorOps::ConditionalSelect(Vector<int>(allow ? -1 : 0), Vector<float>(3.40282347e+38f), speculativeMargin);",
            "    Vector3Wide minExpansion, maxExpansion; BoundingBoxHelpers::GetBoundsExpansion(velocities.Linear, dtWide, angularBoundsExpansion, minExpansion, maxExpansion);",
            "    Vector3Wide negated; negated.X = -maximumBoundsExpansion; negated.Y = -maximumBoundsExpansion; negated.Z = -maximumBoundsExpansion;",
            "    minExpansion.X = VectorOps::Max(negated.X, minExpansion.X); minExpansion.Y = VectorOps::Max(negated.Y, minExpansion.Y); minExpansion.Z = VectorOps::Max(negated.Z, minExpansion.Z);",
            "    maxExpansion.X = VectorOps::Min(maximumBoundsExpansion, maxExpansion.X); maxExpansion.Y = VectorOps::Min(maximumBoundsExpansion, maxExpansion.Y); maxExpansion.Z = VectorOps::Min(maximumBoundsExpansion, maxExpansion.Z);",
            "    bundleMin = positions + (bundleMin + minExpansion); bundleMax = positions + (bundleMax + maxExpansion);",
            "    out[0] = bundleMin.X.v; out[1] = bundleMin.Y.v; out[2] = bundleMin.Z.v; out[3] = speculativeMargin.v; out[4] = bundleMax.X.v; out[5] = bundleMax.Y.v; out[6] = bundleMax.Z.v; return 0; }",
            'extern "C" int ref_covered_types(int* ids, int capacity) { static const int k[] = {%s}; int n = (int)(sizeof(k) / sizeof(k[0])); for (int i = 0; i < n && i < capacity; ++i) ids[i] = k[i]; return n; }' % ", ".join(map(str, covered))]
    return "\n".join(out) + "\n"


def main():
    root, header = sys.argv[1], sys.argv[2]
    tr = Transpiler(root)
    tr.load()
    text, skipped = tr.emit()
    os.makedirs(os.path.dirname(header), exist_ok=True)
    open(header, "w").write(text)
    for s in skipped:
        print("note:", s, file=sys.stderr)
    if "--harness" in sys.argv:
        here = os.path.dirname(os.path.abspath(__file__))
        layouts = json.load(open(os.path.join(here, "..", "..", "tests", "golden", "type_layouts.json")))
        open(sys.argv[sys.argv.index("--harness") + 1], "w").write(harness(tr, layouts))
    print("transpiled %d types, %d methods" % (len(tr.types), sum(len(t.methods) for t in tr.types.values())))


if __name__ == "__main__":
    main()
