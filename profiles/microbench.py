import sys
sys.path.insert(0, '.')
from interpolatory_b200.device import DeviceContext
ctx = DeviceContext(64, 64)
names = {0:'VABSDIFF x8', 1:'FADD pairs x8', 2:'8 SAD + 4 FADD pairs', 3:'8 SAD + 8 FADD pairs', 4:'8 SAD + 2 LDS', 5:'8 SAD + 2 IMAD reg', 6:'8 SAD + 2 IMAD imm', 7:'8 SAD + 2 FADD', 8:'8 SAD + 2 IADD', 9:'8 SAD + 2 VIMNMX', 10:'8 SAD + 1 LDS.128', 11:'8 SAD + 4 LDS.32', 12:'FS register pattern, G=11, 16 warps/SM', 13:'FS register pattern, 8 warps/SM', 14:'FS register pattern, 4 warps/SM', 15:'VABSDIFF4.U8.ACC x8 (4 px each)', 16:'8 VABSDIFF4 + 2 LDS'}
base = None
for mode in range(17):
    best = min(ctx.microbench(mode, 8192)[0] for _ in range(3))
    ms, ops = ctx.microbench(mode, 8192)
    ms = min(ms, best)
    steps = 8192 * 148 * 4 * 512 / 32     # warp-steps
    cyc = ms * 1e-3 * 1.965e9 * 148 * 4 / steps   # SMSP cycles per warp-step
    print(f"mode {mode} {names[mode]:24s}: {ms:.3f} ms, {ops/ms/1e9:.2f} Tpxop/s, {cyc:.2f} SMSP-cycles per step")
