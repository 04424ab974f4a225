"""SASS evidence for the tensor-core / TMA / cluster paths (no GPU needed: cuobjdump on the built objects).
    python tools/sass_facts.py > profiles/<tag>_sass_facts.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA.2CTA", "UTCQMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTMALDG.2D.2CTA", "UTMALDG.2D", "UTCBAR.2CTA.MULTICAST", "UTCBAR",
       "UTCATOMSWS.2CTA", "UTCATOMSWS", "LDTM", "UCGABAR_ARV", "STG.E.ENL2.256", "SYNCS.PHASECHK", "HMMA", "REDG", "RED.E.ADD.F32", "VOTE", "REDUX", "SHFL", "STAS", "ATOMS", "LDS.128", "MAPA", "ATOM.E.OR"]


def main():
    print("# SASS mnemonics per kernel (cuobjdump -sass of mv3d_tf_b200/csrc/_obj/*.o, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4, `.2CTA` = cta_group::2, UTMALDG = "
          "cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, LDTM = tcgen05.ld, UCGABAR = "
          "barrier.cluster, STG.256 = st.global.v8.b32, VOTE = __ballot_sync, REDUX = __reduce_or_sync, SHFL = warp shuffles, STAS = st.async (distributed shared memory + complete_tx), MAPA = mapa (peer shared-memory address), ATOM.E.OR on a mapa address = red.shared::cluster.or.  Pair-kernel template arguments: <BN, PASSES, LEAN, WRES, POOL>.  No legacy HMMA (mma.sync) anywhere.\n")
    print("| object | kernel | " + " | ".join(PAT) + " |")
    print("|---|---|" + "---|" * len(PAT))
    for obj in sorted(glob.glob(os.path.join(ROOT, "mv3d_tf_b200", "csrc", "_obj", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        fn, counts = None, collections.OrderedDict()
        for line in out.split("\n"):
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                fn = re.sub(r"\(.*", "", fn.replace("(int)", "").replace("(bool)", "")).replace("void ", "").replace("mv3d::", "")
                counts[fn] = collections.Counter()
                continue
            if fn is None:
                continue
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                op = m.group(1)
                for p in PAT:
                    if op == p or op.startswith(p + "."):
                        counts[fn][p] += 1
                        break
        for fn, c in counts.items():
            if any(c[p] for p in PAT):
                print("| %s | %s | " % (os.path.basename(obj), fn[:60]) + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + " |")


if __name__ == "__main__":
    main()
