"""Development probe (GPU box): kernel-only timing sweep over variants. Not the bench."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paf_baseband2power_b200 import BMF, Baseband2Power, DeviceBuffer, device_info  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nbeam", type=int, default=1)
ap.add_argument("--ndf", type=int, default=8192)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--kernels", default="ldg,tma")
ap.add_argument("--nsplits", default="0")
ap.add_argument("--variants", default="0")
ap.add_argument("--weaves", default="1")
ap.add_argument("--mode", default="exact")
args = ap.parse_args()

print(json.dumps(device_info(0)))
nb, ndf = args.nbeam, args.ndf
per = ndf * BMF.frame_bytes
buf = DeviceBuffer(nb * per)
for b in range(nb):
    buf.synth_fill(ndf, seed=1 + b, first_word=0, mode=1, offset=b * per)
ptrs = [buf.ptr + b * per for b in range(nb)]
out = DeviceBuffer(nb * BMF.nchan * 4)
for kernel in args.kernels.split(","):
  for variant in args.variants.split(","):
   for weave in args.weaves.split(","):
    os.environ["B2P_VARIANT"] = variant
    os.environ["B2P_WEAVE"] = weave
    for ns in [int(x) for x in args.nsplits.split(",")]:
        st = Baseband2Power(kernel=kernel, nbeam=nb, nsplit=ns, mode=args.mode)
        for _ in range(3):
            st.accumulate_device(ptrs, ndf)
            st.finish_device(out)
        st.set_timing(True)
        times = []
        for _ in range(args.reps):
            st.accumulate_device(ptrs, ndf)
            st.finish_device(out)
            ms, n = st.fused_time_ms()
            times.append(ms / n)
        got = out.download().tobytes()
        if "ref_out" not in globals():
            ref_out = got
        match = got == ref_out
        times.sort()
        best, med = times[0], times[len(times) // 2]
        gb = nb * per / 1e9
        print(json.dumps({"kernel": kernel, "variant": variant, "weave": weave, "mode": args.mode, "nsplit": st.nsplit, "nbeam": nb, "ndf": ndf,
                          "match": match, "best_ms": round(best, 4), "median_ms": round(med, 4),
                          "best_GBps": round(gb / best * 1e3, 1), "median_GBps": round(gb / med * 1e3, 1)}),
              flush=True)
        st.close()
