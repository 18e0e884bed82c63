for v in "--sleep-after-first-rank" "--sleep-after-first-rank --spin 0" ; do echo "== $v"; timeout 80 python bench/stress_bucket_protocol.py --epochs 10 $v 2>&1 | tail -2 | cut -c1-330; done
