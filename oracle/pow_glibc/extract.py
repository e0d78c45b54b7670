import struct, sys
f = open('/lib/x86_64-linux-gnu/libm.so.6','rb').read()
# ELF64 program headers
e_phoff = struct.unpack_from('<Q', f, 0x20)[0]; e_phentsize, e_phnum = struct.unpack_from('<HH', f, 0x36)
loads=[]
for i in range(e_phnum):
    p_type, p_flags, p_offset, p_vaddr, p_paddr, p_filesz, p_memsz, p_align = struct.unpack_from('<IIQQQQQQ', f, e_phoff + i*e_phentsize)
    if p_type == 1: loads.append((p_vaddr, p_offset, p_filesz))
def rd(vaddr, n):
    for va, off, sz in loads:
        if va <= vaddr < va + sz: return f[off + vaddr - va: off + vaddr - va + n]
    raise KeyError(hex(vaddr))
def d(vaddr): return struct.unpack('<d', rd(vaddr, 8))[0]
def u(vaddr): return struct.unpack('<Q', rd(vaddr, 8))[0]
print('neg1?', d(0x99360))
print('ln2hi', d(0xb6b80).hex(), 'ln2lo', d(0xb6b88).hex())
print('A', [d(0xb6b90 + 8*i).hex() for i in range(7)])
print('T0', [d(0xb6bc8 + 8*i).hex() for i in range(4)], 'T127', [d(0xb6bc8 + 32*127 + 8*i).hex() for i in range(4)])
print('exp: invln2N', d(0xb4980).hex(), 'shift', d(0xb4988).hex(), 'negln2hiN', d(0xb4990).hex(), 'negln2loN', d(0xb4998).hex())
print('C', [d(0xb49a0 + 8*i).hex() for i in range(4)])
print('tab0', hex(u(0xb4a30)), hex(u(0xb4a38)), hex(u(0xb4a30+16)), hex(u(0xb4a38+16)))
if len(sys.argv) > 1:
    out = open(sys.argv[1], 'w')
    out.write('/* Constants and tables of glibc 2.39 libm pow (x86_64, FMA variant __pow_fma), read out of /lib/x86_64-linux-gnu/libm.so.6 by\n   oracle/pow_glibc/extract.py: __pow_log_data (ln2hi, ln2lo, poly[7], tab[128] = invc, pad, logc, logctail) and\n   __exp_data (invln2N, shift, negln2hiN, negln2loN, poly[4], tab[256]) as raw IEEE-754 bit patterns. */\n#ifndef SS_POW_TABLE_QUAL\n#define SS_POW_TABLE_QUAL static const\n#endif\n')
    def arr(name, vals):
        out.write('SS_POW_TABLE_QUAL unsigned long long %s[%d] = {\n' % (name, len(vals)))
        for i in range(0, len(vals), 4):
            out.write('    ' + ', '.join('0x%016xull' % v for v in vals[i:i+4]) + ',\n')
        out.write('};\n')
    arr('SS_POW_LOG_HDR', [u(0xb6b80 + 8*i) for i in range(9)])          # ln2hi, ln2lo, A0..A6
    arr('SS_POW_LOG_TAB', [u(0xb6bc8 + 32*i + 8*j) for i in range(128) for j in (0, 2, 3)])   # invc, logc, logctail
    arr('SS_POW_EXP_HDR', [u(0xb4980 + 8*i) for i in range(8)])          # invln2N, shift, negln2hiN, negln2loN, C2..C5
    arr('SS_POW_EXP_TAB', [u(0xb4a30 + 8*i) for i in range(256)])        # tail bits, value bits per entry
    out.close()
