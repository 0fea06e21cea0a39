"""Regenerate profiles/sass/<kernel>.sass from the built libsph_cuda.so (cuobjdump -sass; no GPU needed).
Template instantiations of one kernel go into the same file."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gmu-water-simulation_b200", "libsph_cuda.so")
out_dir = os.path.join(ROOT, "profiles", "sass")
KERNELS = ["k_cell_key_hist", "k_scan", "k_bucket", "k_rank_scatter", "k_density_mask", "k_forces_mask", "k_forces_overflow",
           "k_integrate_collide", "k_density", "k_forces", "k_slab_pack"]
text = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
chunks = re.split(r"(?=^\t\tFunction : )", text, flags=re.M)[1:]
files = collections.defaultdict(list)
for ch in chunks:
    mangled = ch.split("\n", 1)[0].split(":", 1)[1].strip()
    demangled = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
    m = re.search(r"sph::(k_[a-z0-9_]+)", demangled)
    if m and m.group(1) in KERNELS:
        body = ch.rstrip("\n").rsplit("\t\t.....", 1)[0]
        # drop the hex encodings (trailing column and the encoding-only continuation lines): half the size, same listing
        lines = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln) for ln in body.split("\n")]
        files[m.group(1)].append("\n".join(ln for ln in lines if ln.strip()))
os.makedirs(out_dir, exist_ok=True)
for name, parts in files.items():
    with open(os.path.join(out_dir, name + ".sass"), "w") as f:
        f.write("\n".join(parts) + "\n")
    ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", "\n".join(parts), flags=re.M))
    top = ", ".join(f"{k} {v}" for k, v in ops.most_common(6))
    print(f"{name:22s} {len(parts)} instantiation(s), {sum(ops.values()):5d} instructions; {top}")
