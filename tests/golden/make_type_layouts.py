"""Generates tests/golden/type_layouts.json from the reference's own C# sources (run in the build container, where /root/reference exists):

    python tests/golden/make_type_layouts.py [/root/reference]

For every constraint type processor the reference declares (`class XTypeProcessor : {One,Two,Three,Four}Body[Contact]TypeProcessor<TPrestep,
TAccumulatedImpulses, ...> { public const int BatchTypeId = N; }`) it records the body count and the FLATTENED, ORDERED scalar field lists of the
prestep and accumulated-impulse structs (one entry per `Vector<float>` lane row of the AOSOA layout, e.g. "Contact0.OffsetA.X"). This is the layout
`DemoTests/ConstraintDescriptionMappingTests.cs` round-trips in the reference; tests/test_type_layouts.py pins the oracle / device registries and the
row orders their code assumes to it. It also records every type's body access filters (the `Access*` generic arguments of the processor
declaration: WarmStart filters per body, then Solve filters per body) and what each filter gathers / scatters in floats
(`Constraints/IBodyAccessFilter.cs`): the inputs of the SURVEY.md §8d algorithmic-bytes model that `roofline.achieved` is built on.
Parsing only: nothing of the reference is copied into the repository except these field and filter names."""
import json
import os
import re
import sys

ROOT = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = [os.path.join(ROOT, "BepuPhysics", "Constraints"), os.path.join(ROOT, "BepuPhysics", "Constraints", "Contact")]

LEAVES = {  # BepuUtilities wide types: scalar rows in declaration order
    "Vector<float>": [""], "Vector<int>": [""],
    "Vector2Wide": ["X", "Y"], "Vector4Wide": ["X", "Y", "Z", "W"], "Vector3Wide": ["X", "Y", "Z"], "QuaternionWide": ["X", "Y", "Z", "W"],
}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def load_structs():
    structs = {}
    texts = []
    for d in SRC:
        for f in sorted(os.listdir(d)):
            if f.endswith(".cs"):
                texts.append(strip_comments(open(os.path.join(d, f), encoding="utf-8-sig").read()))
    for text in texts:
        for m in re.finditer(r"public\s+(?:unsafe\s+)?struct\s+(\w+)(?:<[^>]*>)?[^{;]*\{", text):
            name = m.group(1)
            # body = up to the matching brace
            depth, i = 1, m.end()
            while depth and i < len(text):
                depth += {"{": 1, "}": -1}.get(text[i], 0)
                i += 1
            body = text[m.end():i - 1]
            # instance fields only: "public T Name;" at struct top level (depth 0 of the body)
            fields, depth, stmt = [], 0, ""
            for ch in body:
                if ch == "{":
                    depth += 1
                    stmt = ""
                elif ch == "}":
                    depth -= 1
                    stmt = ""
                elif depth == 0:
                    if ch == ";":
                        fm = re.match(r"\s*(?:\[[^\]]*\]\s*)*public\s+(?!static|const)([\w<>]+)\s+(\w+)\s*$", stmt.strip(), flags=re.S)
                        if fm:
                            fields.append((fm.group(1), fm.group(2)))
                        stmt = ""
                    else:
                        stmt += ch
            if fields and name not in structs:
                structs[name] = fields
    return structs, texts


def flatten(type_name, structs, prefix=""):
    if type_name in LEAVES:
        return [(prefix + ("." if prefix and leaf else "") + leaf) for leaf in LEAVES[type_name]]
    if type_name not in structs:
        raise KeyError("unknown type %s (prefix %s)" % (type_name, prefix))
    out = []
    for t, n in structs[type_name]:
        out += flatten(t, structs, (prefix + "." if prefix else "") + n)
    return out


def load_filters(texts):
    """Access filter -> floats gathered (read) and velocity floats scattered (written) per body."""
    floats = {"GatherPosition": 3, "GatherOrientation": 4, "GatherMass": 1, "GatherInertiaTensor": 6, "AccessLinearVelocity": 3, "AccessAngularVelocity": 3}
    filters = {}
    for text in texts:
        for m in re.finditer(r"struct\s+(Access\w+)\s*:\s*IBodyAccessFilter\s*\{(.*?)\n\s*\}", text, flags=re.S):
            flags = dict(re.findall(r"public\s+bool\s+(\w+)\s*=>\s*(true|false)", m.group(2)))
            assert set(flags) == set(floats), (m.group(1), flags)
            filters[m.group(1)] = {"read": sum(floats[k] for k, v in flags.items() if v == "true"),
                                   "write": sum(floats[k] for k, v in flags.items() if v == "true" and k.startswith("Access"))}
    return filters


def contact_base_filters(texts):
    """{One,Two}BodyContactTypeProcessor fix their filters in their own base declaration."""
    out = {}
    for text in texts:
        for m in re.finditer(r"class\s+(One|Two)BodyContactTypeProcessor\s*<[^>]*>\s*:\s*\w+TypeProcessor\s*<([^>]*)>", text):
            out[m.group(1)] = [a.strip() for a in m.group(2).split(",") if a.strip().startswith("Access")]
    return out


def main():
    structs, texts = load_structs()
    top = os.path.join(ROOT, "BepuPhysics", "Constraints")
    texts = texts + [strip_comments(open(os.path.join(top, f), encoding="utf-8-sig").read()) for f in ("IBodyAccessFilter.cs",) if os.path.join(top, f) not in SRC]
    filters = load_filters(texts)
    contact_filters = contact_base_filters(texts)
    bodies_of = {"OneBody": 1, "TwoBody": 2, "ThreeBody": 3, "FourBody": 4}
    types = {}
    for text in texts:
        for m in re.finditer(r"class\s+(\w+TypeProcessor)\s*:\s*(One|Two|Three|Four)Body(?:Contact)?TypeProcessor\s*<\s*([\w<>]+)\s*,\s*([\w<>]+)\s*,", text):
            cls, nb, prestep, impulses = m.group(1), m.group(2), m.group(3), m.group(4)
            tail = text[m.end():m.end() + 1500]
            idm = re.search(r"BatchTypeId\s*=\s*(\d+)", tail)
            if not idm:
                continue
            type_id = int(idm.group(1))
            args = re.match(r"[^>{]*", text[m.end():]).group(0)  # the rest of the generic argument list
            access = [a.strip() for a in args.split(",") if a.strip().startswith("Access")]
            if not access:
                access = contact_filters[nb]
            n = bodies_of[nb + "Body"]
            assert len(access) == 2 * n and all(a in filters for a in access), (cls, access)
            types[type_id] = {"processor": cls, "bodies": n, "prestep_struct": prestep, "impulse_struct": impulses,
                              "prestep_rows": flatten(prestep, structs), "impulse_rows": flatten(impulses, structs),
                              "warm_start_filters": access[:n], "solve_filters": access[n:]}
    out = {"generated_from": "BepuPhysics/Constraints/**/*.cs of the reference (bepu/bepuphysics2) by tests/golden/make_type_layouts.py", "access_filters": filters, "types": {str(k): types[k] for k in sorted(types)}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "type_layouts.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %s: %d types" % (path, len(types)))
    for k in sorted(types):
        t = types[k]
        print("%2d %-40s bodies %d prestep %2d impulses %2d" % (k, t["processor"], t["bodies"], len(t["prestep_rows"]), len(t["impulse_rows"])))


if __name__ == "__main__":
    main()
