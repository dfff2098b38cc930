cd $GRAFT_REPO_ROOT
cp soc_project_stereo_matching_b200/lib/libsgm_b200.so /tmp/keep.so
for v in ${VARIANTS:-cur nosleep sb16 sb16ns sb8ns bs16}; do
  cp scripts/micro/libs/$v.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
  echo "== $v"
  python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair
w,h,d=1242,375,128
l,r,_=make_pair(w,h,d,seed=0xB200,texture="noise")
opt=sgm.default_option(max_disparity=d)
with sgm.Context(0) as c:
    c.set_pipeline(sgm.PIPE_REFERENCE); c.configure(w,h,opt)
    ts=[]
    for i in range(30):
        out=c.match(l,r); ts.append(c.last_device_ms())
    c.set_pipeline(sgm.PIPE_HOTPATH); c.configure(w,h,opt)
    th=[]
    for i in range(30):
        c.match(l,r); th.append(c.last_device_ms())
    print("full %.4f hot %.4f diff %.4f ms  checksum %r" % (np.median(ts), np.median(th), np.median(ts)-np.median(th), float(np.nansum(np.where(np.isfinite(out),out,0)))))
PY
done
cp /tmp/keep.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
