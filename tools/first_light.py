"""First-light check on a GPU box: CUDA path vs O-cpu vs O-gpu on config C1 (64^3 ML, 512^2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dvr_harness as H
from visrtx_b200 import capi

def main():
    for rate in (0.125, 0.5, 1.0):
        sc = H.default_scene(64, 512, 512, rate=rate)
        t0 = time.time(); o = H.render_oracle(sc); t1 = time.time()
        c = H.render_cuda(sc)
        cs = H.render_cuda(sc, skip=True)
        r = H.render_refgpu(sc) if H.ob.have_ref_gpu() else None
        print(f"rate {rate}: oracle {t1-t0:.2f}s")
        print("  cuda vs O-cpu  (max/255, psnr):", H.compare_color(c["color"], o["color"], sc.fmt))
        print("  skip vs noskip identical:", np.array_equal(c["color"], cs["color"]), np.array_equal(c["accum"], cs["accum"]))
        if r is not None:
            print("  cuda vs O-gpu  :", H.compare_color(c["color"], r["color"], sc.fmt))
            print("  O-cpu vs O-gpu :", H.compare_color(o["color"], r["color"], sc.fmt))
            print("  accum max abs diff cuda/O-gpu:", float(np.abs(c["accum"]-r["accum"]).max()), " bit-identical frac:", float((c["accum"]==r["accum"]).mean()))
            print("  depth equal cuda/O-gpu:", np.array_equal(c["depth"], r["depth"]), " ids:", np.array_equal(c["objId"], r["objId"]), np.array_equal(c["instId"], r["instId"]), np.array_equal(c["primId"], r["primId"]))
        print("  depth cuda vs O-cpu max diff:", float(np.abs(c["depth"]-o["depth"]).max()))
        hit = (c["depth"] < 1e29).mean()
        print("  hit fraction:", hit)
    np.save("gpurun_out/first_light_color.npy", c["color"].reshape(512,512))

if __name__ == "__main__":
    main()
