# round 2, experiment F: two half-group CTAs per SM (LN=2, shared memory capped), stencil-on-load on/off, per-thread combining loads
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 1500 python tools/sweep.py C4 base "noldsten:B2_NOLDSTEN=1" "ldthreads:B2_LDTHREADS=1" "ldthr_nosten:B2_LDTHREADS=1,B2_NOLDSTEN=1" "ln2:B2_LN=2" "ln2cap:B2_LN=2,B2_SMEMCAP=113" "ln2cap_nosten:B2_LN=2,B2_SMEMCAP=113,B2_NOLDSTEN=1" "ln2cap_ldthr:B2_LN=2,B2_SMEMCAP=113,B2_LDTHREADS=1" "ln2cap_ldthr_nosten:B2_LN=2,B2_SMEMCAP=113,B2_LDTHREADS=1,B2_NOLDSTEN=1"
B2_LN=2 B2_SMEMCAP=113 timeout 100 python tools/copyprobe.py 4097
