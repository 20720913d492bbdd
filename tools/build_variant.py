"""Build a variant of libadyolo_b200.so with extra nvcc defines into build/variants/<name>.so (for A/B runs:
ADYOLO_LIB=build/variants/<name>.so python tools/prof_frontend.py).  Usage: build_variant.py name -DADY_FE2_NT=192 ..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ad-yolo_b200", "csrc")
name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants"); obj_dir = os.path.join(out_dir, name + "_obj")
os.makedirs(obj_dir, exist_ok=True)
srcs = ["fe2.cu", "frontend.cu", "frontend_aux.cu", "assign.cu", "labels.cu", "nms.cu", "gcc_tc.cu", "augment.cu", "tables.cu", "scaler.cu", "api.cu"]
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] + defs
def cc(s):
    o = os.path.join(obj_dir, s.replace(".cu", ".o"))
    subprocess.run(["nvcc", *flags, "-c", os.path.join(CSRC, s), "-o", o], check=True)
    return o
with ThreadPoolExecutor(len(srcs)) as ex:
    objs = list(ex.map(cc, srcs))
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(out_dir, name + ".so"), *objs], check=True)
print(os.path.join(out_dir, name + ".so"))
