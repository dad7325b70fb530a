echo "=== baseline"; NCP=1 LO=40 HI=56 BRIEF= python tools/probe_timeline.py 2>&1 | sed -n 1,20p | cut -c1-150
echo "=== no proxy fence (timing only)"; DBG=512 NCP=1 LO=40 HI=56 python tools/probe_timeline.py 2>&1 | sed -n 1,20p | cut -c1-150
