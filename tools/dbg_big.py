import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import corpus
import lzma_rs_b200 as L
ctx = L.Context()
for n in (3000, 60000):
    data = corpus.mixed_text(4242, n)
    for lc in (3, 0, 8):
        blob = corpus.lzma_alone_known_size(data, dict_size=1 << 16, lc=lc)
        r = ctx.decompress_one(0, blob)
        print("n", n, "lc", lc, "ok", r.ok, r.display, len(r.data), r.data == data, flush=True)
