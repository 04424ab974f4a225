"""TEST INFRASTRUCTURE ONLY.  Compiles oracle/oracle_c.c -> oracle/_build/liboracle_c.so (gcc)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle_c.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "oracle_c.c")
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(
        ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-o", OUT, src, "-lm"],
        check=True,
    )
    return OUT


if __name__ == "__main__":
    print(build(force=True))
