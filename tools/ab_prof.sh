cd /root/repo
python tools/ab_run.py cfg2s > /dev/null 2>&1   # links the variant libraries into /tmp
for v in relay0 relay1; do echo "#### $v"; SWN_B200_LIB=/tmp/libswn_$v.so python tools/gpu_prof_core.py cfg2s 2>&1 | grep -v "variant=1"; done
