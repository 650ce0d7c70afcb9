import sys, importlib
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
for k in range(6):
    ctx.scan_to_pose(synth.scan("HDL-64", k))
