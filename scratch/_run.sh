bash scratch/qb.sh "--workload encoder6"
for d in 6000 12000 18000; do echo "== stagger $d"; MSDA_B200_LIB=richsem_b200/lib/variants/libmsda_stg$d.so bash scratch/qb.sh "--workload encoder6"; done
