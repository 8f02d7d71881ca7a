"""Static cross-check of the hand-written Fortran binding against the C header (CPU).

The build image has no Fortran compiler, so ``metalquicha_b200/fortran/mqc_b200_iface.f90`` cannot be
compiled here.  What can go wrong in a hand-written ``bind(C)`` interface is mechanical -- a missing or
extra argument, two arguments swapped, ``value`` forgotten on a scalar (the callee would then receive an
address) or added to an array, a 32-bit kind where the C side reads 64 bits -- and all of that is visible
by reading both files.  This test parses every prototype of ``include/mqcb200.h`` and every interface body
of the module and compares them argument by argument; it also checks that the wrapper module and the
refusing stub only use what the interface module exports.  (The reference's own bindings follow the same
pattern and are the model: backends/cuest/bindings/cublas.f90:20-44.)
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mqcb200.h")
FDIR = os.path.join(ROOT, "metalquicha_b200", "fortran")
IFACE = os.path.join(FDIR, "mqc_b200_iface.f90")


# ---- the C side ------------------------------------------------------------------------------------
def c_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"\b(int|void|const\s+char\s*\*)\s+(mqcb200_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                arr = re.search(r"\[\s*\w*\s*\]\s*$", a)           # `char id[128]`, `double ms[N]`: pointers
                if arr:
                    a = a[:arr.start()].strip()
                pm = re.match(r"(.*?)(\w+)$", a)
                ctype, pname = pm.group(1).strip(), pm.group(2)
                ctype = re.sub(r"\s*\*\s*", "*", ctype)
                if arr:
                    ctype += "*"
                params.append((ctype, pname))
        protos[name] = (re.sub(r"\s+", " ", ret), params)
    return protos


def c_class(ctype):
    """-> (base, pointer_depth) with const dropped."""
    depth = ctype.count("*")
    base = ctype.replace("*", "").replace("const", "").strip()
    return base, depth


# ---- the Fortran side ------------------------------------------------------------------------------
def _join_continuations(src):
    out, cur = [], ""
    for raw in src.splitlines():
        line = raw.split("!")[0].rstrip()          # no character literal in the interface bodies holds a '!'
        if not line.strip():
            continue
        if cur:
            line = line.lstrip()
            if line.startswith("&"):
                line = line[1:]
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        out.append(cur + line)
        cur = ""
    return out


def fortran_interfaces():
    lines = _join_continuations(open(IFACE).read())
    faces = {}
    i = 0
    head = re.compile(r"^\s*(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*\"(\w+)\"\s*\)"
                      r"(?:\s*result\s*\(\s*(\w+)\s*\))?", re.I)
    while i < len(lines):
        m = head.match(lines[i])
        if not m:
            i += 1
            continue
        kind, fname, arglist, cname, result = m.group(1).lower(), m.group(2), m.group(3), m.group(4), m.group(5)
        args = [a.strip().lower() for a in arglist.split(",") if a.strip()]
        decls = {}
        i += 1
        while not re.match(r"^\s*end\s+(function|subroutine)", lines[i], re.I):
            dm = re.match(r"^\s*(integer|real|type|character)\s*\(([^)]*)\)\s*((?:,\s*[\w()=]+\s*)*)::\s*(.*)$", lines[i], re.I)
            if dm:
                base = dm.group(1).lower()
                kindspec = dm.group(2).replace(" ", "").lower().replace("kind=", "")
                attrs = [a.strip().lower() for a in dm.group(3).split(",") if a.strip()]
                for ent in re.findall(r"(\w+)\s*(\([^)]*\))?", dm.group(4)):
                    decls[ent[0].lower()] = {"base": base, "kind": kindspec, "value": "value" in attrs,
                                             "array": bool(ent[1]), "attrs": attrs}
            i += 1
        faces[cname] = {"kind": kind, "fname": fname, "args": args, "decls": decls,
                        "result": result.lower() if result else None}
        i += 1
    return faces


SCALAR = {"int": ("integer", {"c_int"}), "double": ("real", {"c_double"}), "size_t": ("integer", {"c_size_t"}),
          "int64_t": ("integer", {"c_int64_t", "c_long_long"}), "uint64_t": ("integer", {"c_int64_t", "c_long_long"}),
          "long long": ("integer", {"c_long_long", "c_int64_t"})}
POINTEE = {"int": ("integer", {"c_int"}), "double": ("real", {"c_double"}), "size_t": ("integer", {"c_size_t"}),
           "int64_t": ("integer", {"c_int64_t", "c_long_long"}), "char": ("character", {"c_char"})}


def _mismatch(ctype, d):
    base, depth = c_class(ctype)
    if depth == 0:
        fb, kinds = SCALAR[base]
        if not d["value"]:
            return "a C scalar passed by value needs the VALUE attribute"
        if d["array"] or d["base"] != fb or d["kind"] not in kinds:
            return f"expected {fb}({'|'.join(sorted(kinds))}), value"
        return None
    # pointers
    if d["base"] == "type" and d["kind"] == "c_ptr":
        if base == "void" and depth == 2:
            return "void** must be a c_ptr passed by reference (no VALUE)" if d["value"] else None
        return None if d["value"] else "a pointer handed over as type(c_ptr) must carry VALUE"
    if d["value"]:
        return "a pointer argument declared as a typed dummy must not carry VALUE"
    if base == "void" or depth != 1:
        return "void* / multi-level pointers must be type(c_ptr)"
    fb, kinds = POINTEE[base]
    if d["base"] != fb or d["kind"] not in kinds:
        return f"expected {fb}({'|'.join(sorted(kinds))}) by reference"
    return None


def test_every_prototype_has_a_matching_interface_body():
    protos, faces = c_prototypes(), fortran_interfaces()
    assert len(protos) >= 49, sorted(protos)                       # the whole ABI surface was parsed
    assert set(faces) == set(protos), (sorted(set(protos) - set(faces)), sorted(set(faces) - set(protos)))
    problems = []
    for name, (ret, params) in sorted(protos.items()):
        f = faces[name]
        if f["fname"].lower() != name:
            problems.append(f"{name}: Fortran name {f['fname']} differs from the bound name")
        if (ret == "void") != (f["kind"] == "subroutine"):
            problems.append(f"{name}: C returns {ret}, Fortran declares a {f['kind']}")
        if f["kind"] == "function":
            r = f["decls"].get(f["result"] or name)
            if ret == "int" and not (r and r["base"] == "integer" and r["kind"] == "c_int" and not r["array"]):
                problems.append(f"{name}: the status result must be integer(c_int)")
        if len(params) != len(f["args"]):
            problems.append(f"{name}: {len(params)} C parameters, {len(f['args'])} Fortran dummies")
            continue
        for pos, ((ctype, pname), dummy) in enumerate(zip(params, f["args"])):
            d = f["decls"].get(dummy)
            if d is None:
                problems.append(f"{name}: dummy '{dummy}' is not declared")
                continue
            why = _mismatch(ctype, d)
            if why:
                problems.append(f"{name}: argument {pos + 1} ({ctype} {pname} <-> {dummy}): {why}")
    assert not problems, "\n".join(problems)


def test_argument_names_follow_the_header_order():
    """Same names in the same order (the interface bodies were written from the header): a swap of two
    arguments of the same type, which the type check above cannot see, shows up here."""
    protos, faces = c_prototypes(), fortran_interfaces()
    alias = {"id": {"id", "unique_id"}}
    swapped = []
    for name, (_, params) in sorted(protos.items()):
        c_names = [p.lower() for _, p in params]
        f_names = faces[name]["args"]
        if sorted(c_names) == sorted(f_names) and c_names != f_names:
            swapped.append(f"{name}: C order {c_names}, Fortran order {f_names}")
        for c, f in zip(c_names, f_names):
            if c != f and f in c_names and c in f_names and f not in alias.get(c, ()):
                swapped.append(f"{name}: '{c}' and '{f}' trade places")
    assert not swapped, "\n".join(sorted(set(swapped)))


def test_public_list_and_users_of_the_interface_module():
    src = "\n".join(_join_continuations(open(IFACE).read()))
    public = set()
    for m in re.finditer(r"^\s*public\s*::\s*(.*)$", src, re.M | re.I):
        public.update(x.strip().lower() for x in m.group(1).split(","))
    faces = fortran_interfaces()
    missing = [n for n in faces if n.lower() not in public]
    assert not missing, f"interface bodies not in the PUBLIC list: {missing}"
    consts = {m.group(1).lower() for m in re.finditer(r"parameter\s*::\s*(\w+)", src, re.I)}
    exported = public
    assert {n.lower() for n in faces} | consts >= exported, sorted(exported - ({n.lower() for n in faces} | consts))
    # the wrapper module and the stub import only what exists
    for fn in os.listdir(FDIR):
        if fn == os.path.basename(IFACE) or not fn.endswith(".f90"):
            continue
        text = "\n".join(_join_continuations(open(os.path.join(FDIR, fn)).read()))
        for m in re.finditer(r"use\s+mqc_b200_iface\s*,\s*only\s*:\s*(.*)$", text, re.M | re.I):
            for item in m.group(1).split(","):
                item = item.split("=>")[-1].strip().lower()
                assert item in exported, f"{fn} imports '{item}', which mqc_b200_iface does not export"


def test_constants_agree_with_the_header():
    header = open(HEADER).read()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(MQCB200_\w+)\s+(-?\d+)\b", header)}
    src = "\n".join(_join_continuations(open(IFACE).read()))
    params = {m.group(1).upper(): int(m.group(2)) for m in
              re.finditer(r"integer\(c_int\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(-?\d+)", src, re.I)}
    assert params, "no integer parameters found in the interface module"
    for name, value in params.items():
        assert name in defines, f"{name} is not a #define of the header"
        assert defines[name] == value, f"{name}: header {defines[name]}, Fortran {value}"
    for must in ("MQCB200_OK", "MQCB200_FAIL", "MQCB200_BAD_HANDLE", "MQCB200_NUM_TIMERS"):
        assert must in params


def _call_argument_counts(text):
    """Every ``mqcb200_xxx( ... )`` reference in joined Fortran source -> (name, number of actual arguments)."""
    found = []
    for m in re.finditer(r"\b(mqcb200_\w+)\s*\(", text):
        depth, i, args, cur = 1, m.end(), 0, False
        while i < len(text) and depth:
            ch = text[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                args += 1
            elif not ch.isspace():
                cur = True
            i += 1
        found.append((m.group(1), args + 1 if cur else 0))
    return found


def test_wrapper_calls_pass_as_many_arguments_as_the_prototypes_take():
    """The wrapper module (mqc_b200_fock.f90: the routines a maintainer calls in place of build_fock_df,
    build_df_tensor, response_operator_df ...) must hand every entry point exactly the arguments the header
    lists -- with explicit interfaces a compiler would refuse anything else; here the test does."""
    protos = c_prototypes()
    n_calls = 0
    for fn in sorted(os.listdir(FDIR)):
        if not fn.endswith(".f90") or fn == os.path.basename(IFACE):
            continue
        text = "\n".join(_join_continuations(open(os.path.join(FDIR, fn)).read()))
        for name, n_args in _call_argument_counts(text):
            assert name in protos, f"{fn} calls {name}, which the header does not declare"
            assert n_args == len(protos[name][1]), f"{fn}: {name} called with {n_args} arguments, the header takes {len(protos[name][1])}"
            n_calls += 1
    assert n_calls >= 15            # the wrapper really goes through the interface module


def _module_procedures(path):
    """name -> dummy-argument list of every subroutine/function of a module file (joined source)."""
    text = _join_continuations(open(path).read())
    procs = {}
    for line in text:
        m = re.match(r"^\s*(?:pure\s+)?(subroutine|function)\s+(\w+)\s*\(([^)]*)\)", line, re.I)
        if m:
            procs[m.group(2).lower()] = [a.strip().lower() for a in m.group(3).split(",") if a.strip()]
    return procs


def test_stub_module_offers_the_same_entry_points_with_the_same_arguments():
    """The refusing stub (what a build without the engine compiles, the reference's own arrangement:
    src/methods/mqc_libcint_bridge_stub.f90) must be call-compatible with the real module: same public names,
    same dummy arguments in the same order -- otherwise a call site compiles against one and not the other."""
    real = os.path.join(FDIR, "mqc_b200_fock.f90")
    stub = os.path.join(FDIR, "mqc_b200_fock_stub.f90")

    def public_names(path):
        names = set()
        for line in _join_continuations(open(path).read()):
            m = re.match(r"^\s*public\s*::\s*(.*)$", line, re.I)
            if m:
                names.update(x.strip().lower() for x in m.group(1).split(","))
        return names
    pub_real, pub_stub = public_names(real), public_names(stub)
    assert pub_real == pub_stub, (sorted(pub_real - pub_stub), sorted(pub_stub - pub_real))
    procs_real, procs_stub = _module_procedures(real), _module_procedures(stub)
    for name in sorted(pub_real):
        assert name in procs_real and name in procs_stub, name
        assert procs_real[name] == procs_stub[name], f"{name}: {procs_real[name]} vs {procs_stub[name]}"
