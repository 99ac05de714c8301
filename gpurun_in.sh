for g in 4 20; do
echo "=== GROUP $g"
HDB_BULK_GROUP=$g PROBE_NOCLK=1 PROBE_M=8192 timeout 100 python tools/phase_probe.py 2>&1 | grep "kernel ms"
done
echo "=== with clocks GROUP 4, none / disp only"
HDB_BULK_GROUP=4 PROBE_ONLY="none,disp only,react only,forces only" PROBE_M=8192 timeout 100 python tools/phase_probe.py 2>&1 | grep -v "^---" | tail -24
