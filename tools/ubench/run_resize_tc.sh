cd tools/ubench
for a in "4 640 480 8 2" "3 752 480 8 3" "5 333 241 5 5" "1024 640 480 8 32" "1024 640 480 8 64" "1024 752 480 8 32"; do
  echo "== $a"; timeout 90 ./resize_tc $a 2>&1 | tail -14
done
