for sh in 28 27 26; do
ncu --kernel-name regex:"qprobe" --launch-skip 3 -c 1 --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_op_read_lookup_miss.sum,lts__t_sectors_op_write_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/hit_$sh.csv python tools/prof_qpart.py 5e8 134217728 2 $sh > /dev/null 2>&1
echo "== shift $sh"; grep -a "qprobe" gpurun_out/hit_$sh.csv | awk -F'","' '{print $(NF-2), $NF}'
done
