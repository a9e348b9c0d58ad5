import sys, os, io
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import corpus, oracle_py as oracle
import lzma_rs_b200 as L
from lzma_rs_b200 import _native
ctx = L.Context()
data = corpus.mixed_text(4242, 3000)
for lc in (3, 0):
    blob = corpus.lzma_alone_known_size(data, dict_size=1 << 16, lc=lc)
    h = ctx.raw_new(0, lc, 0, 2, 1 << 16)
    opts = L.decompress.Options(L.decompress.UnpackedSize.UseProvided(len(data)))
    r = h.decompress(blob[13:], opts)
    print("lc", lc, "raw:", r.ok, r.display, len(r.data), r.consumed, r.data[:80] == data[:80], flush=True)
    k = 0
    while k < min(len(r.data), len(data)) and r.data[k] == data[k]:
        k += 1
    print("  first mismatch at", k, flush=True)
    r2 = ctx.decompress_one(0, blob)
    print("  one-shot:", r2.ok, len(r2.data), flush=True)
