#!/bin/bash
# compute-sanitizer memcheck + racecheck (+ synccheck) over __graft_entry__.smoke(): every extractor kernel, the all-pairs search and
# SearchForInitialization on one 752x480 frame pair.  Logs go to gpurun_out/ (copy the summaries into profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
    timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
    tail -4 gpurun_out/sanitizer_$tool.log
done
