"""The JNI glue cannot be linked here (no JDK in the image), but it can be compiled: java/jni/sieve_jni.c is built with
-DSIEVE_HAVE_JNI against tests/jni_stub/jni.h (the handful of JNI declarations it uses, signatures of the JNI specification)
and include/sieve_b200.h, warnings as errors.  Then every `native` method of SieveB200.java must have its
Java_beast_evolution_likelihood_SieveB200_<name> export and every export a native declaration, with matching arity."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "beast", "evolution", "likelihood", "SieveB200.java")
GLUE = os.path.join(ROOT, "java", "jni", "sieve_jni.c")


def _native_methods():
    src = open(JAVA).read()
    out = {}
    for m in re.finditer(r"public\s+static\s+native\s+\w+(?:\[\])?\s+(\w+)\s*\(([^)]*)\)\s*;", src, re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def test_glue_compiles_and_matches_the_native_declarations(tmp_path):
    obj = str(tmp_path / "sieve_jni.o")
    cmd = ["gcc", "-std=c11", "-DSIEVE_HAVE_JNI", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fPIC",
           "-I" + os.path.join(ROOT, "tests", "jni_stub"), "-I" + os.path.join(ROOT, "include"), "-c", GLUE, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", "--defined-only", "-g", obj], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT Java_beast_evolution_likelihood_SieveB200_(\w+)", syms))
    natives = _native_methods()
    assert len(natives) >= 28
    assert exported == set(natives), (sorted(exported - set(natives)), sorted(set(natives) - exported))
    # arity: JNIEnv*, jclass + the Java parameters
    glue = open(GLUE).read()
    for name, n_args in natives.items():
        m = re.search(r"FN\(%s\)\s*\(([^)]*)\)" % name, glue, re.S)
        assert m, name
        c_args = [a for a in m.group(1).split(",") if a.strip()]
        assert len(c_args) == n_args + 2, (name, len(c_args), n_args)
    # every C function the glue forwards to is declared in the public header
    header = open(os.path.join(ROOT, "include", "sieve_b200.h")).read()
    for fn in set(re.findall(r"\b(sieve_[a-z_0-9]+)\s*\(", glue)):
        assert re.search(r"\b%s\s*\(" % fn, header), fn


def test_java_core_passes_the_model_flags():
    src = open(os.path.join(ROOT, "java", "beast", "evolution", "likelihood", "ScsCudaLikelihoodCore.java")).read()
    assert "singleADO ? SieveB200.FLAG_SINGLE_ADO : 0" in src and "ascBias ? SieveB200.FLAG_CONST_PARTIALS : 0" in src
    flags = dict(re.findall(r"(FLAG_\w+)\s*=\s*(\d+)", open(JAVA).read()))
    from sieve_b200 import _lib
    for name, val in flags.items():
        assert getattr(_lib, name) == int(val), name
