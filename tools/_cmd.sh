for ch in 4 8 12 16 1000; do
LWS_C8_CH=$ch ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --profile-from-start off -k regex:conv3d_c8p_kernel -s 5 -c 1 --csv --log-file gpurun_out/c8p_q36_$ch.csv python tools/profile_step.py --batch 8 --iters 1 > /dev/null 2>&1
echo "CH $ch: $(grep -o '"gpu__time_duration.sum","ns","[0-9,]*"\|"dram__bytes_read.sum","byte","[0-9,]*"\|"dram__bytes_read.sum","Mbyte","[0-9.]*"' gpurun_out/c8p_q36_$ch.csv | tr '\n' ' ')"
done
