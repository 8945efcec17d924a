"""profiles/sass/<tag>_<kernel>.sass: cuobjdump -sass of selected kernels of the in-tree library.

    python scripts/dump_sass.py r02
"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {                                   # mangled-name fragment -> file stem
    "list_kernelILi8ELi48ELb1E": "list_kernel_8_p16", "list_kernelILi2ELi48ELb1E": "list_kernel_2_p16",
    "probe_kernelILi4ELb1E": "probe_kernel_4_p16", "prep_scatter_kernelILb1E": "prep_scatter_kernel_p16",
    "scan_kernelIiLb1ELb1E": "scan_kernel_i32_p16_vec", "heaps_kernelIiE": "heaps_kernel_i32", "grid_kernel": "bernoulli_grid_kernel",
    "9ks_kernelILb1E": "ks_kernel_smem", "coo_count_kernel": "coo_count_kernel", "spectrum_kernel": "spectrum_kernel",
}
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
text = subprocess.run(["cuobjdump", "-sass", os.path.join(REPO, "pangenomix_b200", "libpgx_b200.so")],
                      check=True, capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", text)
os.makedirs(os.path.join(REPO, "profiles", "sass"), exist_ok=True)
for part in parts[1:]:
    name = part.split("\n", 1)[0].strip()
    for fragment, stem in WANT.items():
        if fragment in name:
            body = part.split("\n\t\t.......", 1)[0]
            with open(os.path.join(REPO, "profiles", "sass", "%s_%s.sass" % (tag, stem)), "w") as f:
                f.write("Function : " + body.rstrip() + "\n")
            print(stem, len(body.splitlines()), "lines")
            break
