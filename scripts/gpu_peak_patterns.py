import ctypes
import safe_exploration_b200 as se
lib = se._lib.load()
for n, pat in ((96, 0), (96, 1), (96, 2), (96, 3), (128, 1), (192, 1), (256, 0), (256, 1), (64, 1), (160, 1), (240, 1)):
    t = ctypes.c_double()
    se._lib.check(lib.segp_i8_peak_pattern(0, n, pat, 30000, ctypes.byref(t)))
    clk = 2.0 * 128 * n * 32 * 148 / (t.value * 1e12) * 1.965e9
    print("N=%3d pattern %d: %7.1f TOP/s  (~%.1f clk per MMA at 1965 MHz; nominal %.0f)" % (n, pat, t.value, clk, n / 2))
