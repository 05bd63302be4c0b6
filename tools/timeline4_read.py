"""Offline reader of tools/timeline4.py dumps: aligns SM clocks and prints skew statistics. usage: timeline4_read.py file.npy [strip...]"""
import sys
import numpy as np
G, K = 132, 4
raw = np.load(sys.argv[1])
ph = raw[:G * 16].reshape(G, 16)
tl = raw[G * 16:].reshape(G, 64, 16)
gt0, ck0, gt1, ck1 = (tl[:, 63, i].astype(np.float64) for i in range(4))
rate = (ck1 - ck0) / (gt1 - gt0)                      # cycles per ns
print("clock rate cycles/ns: %.4f .. %.4f" % (rate.min(), rate.max()))
st = tl[:, :32, :].astype(np.int64)
base = int(tl[:, 63, 0].min())
t = ((st - (base & 0xFFFFFFFF)) & 0xFFFFFFFF).astype(np.float64) * 1.965      # globaltimer ns (low 32 bits) -> cycles at 1965 MHz
t[st == 0] = np.nan
c, o = t[:, :, :8], t[:, :, 8:]
hs = np.arange(G) < G - K
T = slice(2, 30)
np.set_printoptions(linewidth=250, suppress=True)
Cend = c[hs][:, :, 4]
print("period (cycles) %.0f" % np.nanmean(np.diff(np.nanmax(Cend, 0))[T]))
print("C end: spread %.0f  mean-to-max %.0f" % (np.nanmean((np.nanmax(Cend, 0) - np.nanmin(Cend, 0))[T]), np.nanmean((np.nanmax(Cend, 0) - np.nanmean(Cend, 0))[T])))
last = np.nanargmax(Cend, 0)[T]
print("last CTA in C per strip:", last)
Pend = c[:, :, 5]
val = np.ones(G, bool); val[0] = False; val[G - K - 1] = False
d = (Pend[val] - np.nanmax(Cend, 0)[None, :])[:, T]
print("poll end - last C end: min %.0f mean %.0f max %.0f" % (np.nanmean(np.nanmin(d, 0)), np.nanmean(d), np.nanmean(np.nanmax(d, 0))))
names = ["pre", "A", "B", "C", "poll", "sum", "send"]
for i, nm in enumerate(names):
    d = (c[hs][:, :, i + 1] - c[hs][:, :, i])[:, T]
    print("crit %-5s mean %5.0f  max-over-CTAs %5.0f" % (nm, np.nanmean(d), np.nanmean(np.nanmax(d, 0))))
for i, nm in enumerate(["G wait", "a", "x3 wait", "corr+send", "V wait", "W", "vb"]):
    d = (o[:, :, i + 1] - o[:, :, i])[:, T]
    print("off  %-9s mean %5.0f  max-over-CTAs %5.0f min %5.0f" % (nm, np.nanmean(d), np.nanmean(np.nanmax(d, 0)), np.nanmean(np.nanmin(d, 0))))
for it in [int(x) for x in sys.argv[2:]]:
    base = np.nanmin(t[:, it, 0])
    print("strip", it, "crit[start pre A B C poll sum send] off[start G a x3 corr V W vb]")
    for g in range(G):
        print("%3d" % g, np.nan_to_num(t[g, it] - base).astype(int))
