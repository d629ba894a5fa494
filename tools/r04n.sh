# host packer: AVX2 + scalar scatter, no prefetch (as shipped so far) vs AVX-512 branch-free + software prefetch
grep -m1 "model name" /proc/cpuinfo; nproc; grep -c avx512f /proc/cpuinfo
echo "== old (avx2, no prefetch)"; MGQP_HOST_PACK_ISA=avx2 MGQP_HOST_PACK_PREFETCH=0 python tools/pack_probe.py 2>&1 | grep -v "flags=32" | tail -11
echo "== avx2 + prefetch"; MGQP_HOST_PACK_ISA=avx2 python tools/pack_probe.py 2>&1 | grep -v "flags=32" | tail -11
echo "== new (avx512 + prefetch 2)"; python tools/pack_probe.py 2>&1 | grep -v "flags=32" | tail -11
echo "== new, prefetch 4"; MGQP_HOST_PACK_PREFETCH=4 python tools/pack_probe.py 2>&1 | grep -v "flags=32" | tail -11
