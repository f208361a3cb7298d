#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (oracle).  Registers the B200 backend inside a COPY of the five reference files INTEGRATION.md
names, so that the reference's own main() can be linked with the plug-in class (make -C oracle b200 ->
oracle/_ref/svmultiphysics_b200) and `<Linear_algebra type="b200">` can be selected from a real solver.xml.

    python patch_reference.py <reference Code/Source/solver dir> <output dir>

The copies are written to <output dir> (under oracle/_ref/, git-ignored); nothing of the reference is committed - the edits are
anchored insertions / replacements described here, each anchor a short unique substring of the file it is applied to.  Every
edit asserts that its anchor occurs exactly as often as expected, so a changed reference fails the build instead of silently
building an unpatched solver.
"""
import os
import sys

INCLUDE = '#include "B200LinearAlgebra.h"\n'
DYN = "dynamic_cast<B200LinearAlgebra*>(com_mod.eq[com_mod.cEq].linear_algebra)"


def sub(text, anchor, new, count=1, where=""):
    n = text.count(anchor)
    assert n == count, f"{where}: anchor {anchor!r} occurs {n} times, expected {count}"
    return text.replace(anchor, new)


def after_last_include(text, line):
    i = text.rfind("#include")
    j = text.index("\n", i) + 1
    return text[:j] + line + text[j:]


def patch_consts(t):
    # the enumerator (consts.h:503-508)
    i = t.index("enum class LinearAlgebraType")
    j = t.index("}", i)
    body = t[i:j]
    assert "b200" not in body and body.rstrip().endswith("trilinos"), "consts.h: unexpected LinearAlgebraType body"
    return t[:i] + body.rstrip() + ",\n  b200\n" + t[j:]


def patch_linear_algebra(t):
    # items 2 and 3: the two name maps and the factory (LinearAlgebra.cpp:36-48, 72-90)
    t = after_last_include(t, INCLUDE)
    t = sub(t, '{"trilinos", consts::LinearAlgebraType::trilinos}',
            '{"trilinos", consts::LinearAlgebraType::trilinos},\n  {"b200", consts::LinearAlgebraType::b200}', where="LinearAlgebra.cpp")
    t = sub(t, '{consts::LinearAlgebraType::trilinos, "trilinos"}',
            '{consts::LinearAlgebraType::trilinos, "trilinos"},\n  {consts::LinearAlgebraType::b200, "b200"}', where="LinearAlgebra.cpp")
    t = sub(t, "interface = new TrilinosLinearAlgebra();\n    break;",
            "interface = new TrilinosLinearAlgebra();\n    break;\n\n    case consts::LinearAlgebraType::b200:\n"
            "      interface = new B200LinearAlgebra();\n    break;", where="LinearAlgebra.cpp")
    return t


def patch_eq_assem(t):
    # item 4: whole-mesh assembly goes to the device when the equation's backend offers it (eq_assem.cpp:398-402);
    # item 5: moving-mesh face vectors reach the device after fsils_bc_update (eq_assem.cpp:370)
    t = after_last_include(t, INCLUDE)
    i = t.index("void global_eq_assem(")
    j = t.index("switch (eq.phys)", i)
    hook = ("if (auto* b200 = dynamic_cast<B200LinearAlgebra*>(eq.linear_algebra)) {\n"
            "    if (b200->assemble_mesh(com_mod, lM, Ag, Yg, Dg, &cep_mod)) return;\n  }\n\n  ")
    t = t[:j] + hook + t[j:]
    t = sub(t, "fsils_bc_update(com_mod.lhs, lBc.lsPtr, lFa.nNo, nsd, sVl);",
            "fsils_bc_update(com_mod.lhs, lBc.lsPtr, lFa.nNo, nsd, sVl);\n"
            f"  if (auto* b200 = {DYN}) b200->update_faces(com_mod);", where="eq_assem.cpp")
    return t


def patch_main(t):
    # item 4 (ustruct): Kd lives on the device (main.cpp:526)
    t = after_last_include(t, INCLUDE)
    t = sub(t, "ustruct::ustruct_r(com_mod, Yg);",
            f"auto* b200 = {DYN};\n        if (!b200 || !b200->ustruct_r(com_mod, Yg)) ustruct::ustruct_r(com_mod, Yg);", where="main.cpp")
    return t


def patch_set_bc(t):
    # item 4 (Neumann faces): set_bc_neu_l (set_bc.cpp:1446-1449)
    t = after_last_include(t, INCLUDE)
    t = sub(t, "eq_assem::b_neu_folw_p(com_mod, lBc, lFa, hg, Dg);",
            f"auto* b200 = {DYN};\n    if (!b200 || !b200->assemble_follower_face(com_mod, lFa, hg, Dg)) eq_assem::b_neu_folw_p(com_mod, lBc, lFa, hg, Dg);",
            where="set_bc.cpp")
    t = sub(t, "eq_assem::b_assem_neu_bc(com_mod, lFa, hg, Yg);",
            f"auto* b200 = {DYN};\n    if (!b200 || !b200->assemble_face(com_mod, lFa, hg, Yg)) eq_assem::b_assem_neu_bc(com_mod, lFa, hg, Yg);",
            where="set_bc.cpp")
    return t


EDITS = {"consts.h": patch_consts, "LinearAlgebra.cpp": patch_linear_algebra, "eq_assem.cpp": patch_eq_assem,
         "main.cpp": patch_main, "set_bc.cpp": patch_set_bc}


def main():
    src, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in EDITS.items():
        with open(os.path.join(src, name)) as f:
            text = f.read()
        new = fn(text)
        assert new != text, name
        path = os.path.join(out, name)
        if not (os.path.exists(path) and open(path).read() == new):     # keep time stamps when nothing changed
            with open(path, "w") as f:
                f.write(new)
    print("patched:", ", ".join(EDITS))


if __name__ == "__main__":
    main()
