import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openmmgridforce_b200 as gf
dev = gf.Device(0)
rt = ctypes.CDLL("libcudart.so.12") if False else None
for mb in (32, 64, 89, 96, 128, 1024):
    a = dev.bench_sector_gather(mb << 20, 1 << 24, 10)
    b = dev.bench_sector_gather(mb << 20, 1 << 24, -10)
    print(f"{mb:5d} MB: LDG.256 {a:8.1f} GB/s = {a/32:7.1f} G loads/s | LDG.128 {b:8.1f} GB/s = {b/16:7.1f} G loads/s")
