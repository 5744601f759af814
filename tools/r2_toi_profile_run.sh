O=gpurun_out/r3toi; mkdir -p $O
python tools/ncu_toi_target.py settled 16384 > $O/settled_plain.txt 2>&1
python tools/ncu_toi_target.py landing 16384 45 > $O/landing_plain.txt 2>&1
for m in settled landing; do
  skip=202; [ $m = landing ] && skip=45
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:toi_kernel --launch-skip $skip --launch-count 1 -o $O/$m python tools/ncu_toi_target.py $m 16384 45 > $O/$m.log 2>&1
  ncu -i $O/$m.ncu-rep --page source --csv --print-source sass > $O/${m}_src.csv 2>> $O/$m.log
  python tools/ncu_summary.py $O/$m.ncu-rep > $O/${m}_summary.txt 2>> $O/$m.log
  python tools/ncu_hot.py $O/${m}_src.csv 45 > $O/${m}_hot.txt 2>> $O/$m.log
  ncu -i $O/$m.ncu-rep --page source --csv --print-source cuda > $O/${m}_cuda.csv 2>> $O/$m.log
  rm -f $O/$m.ncu-rep $O/${m}_src.csv
done
cat $O/settled_plain.txt $O/landing_plain.txt | tail -4
