#!/bin/bash
# run an experiment binary that prints "RESTART <idx>" when a configuration faulted; continue after the faulting one
i=0
while [ "$i" != "-1" ]; do
  timeout 120 "$1" $i > /tmp/exp_out.txt 2>&1
  grep -v RESTART /tmp/exp_out.txt
  n=$(grep RESTART /tmp/exp_out.txt | tail -1 | awk '{print $2}')
  if [ -z "$n" ]; then echo "no RESTART marker (crash/timeout) after index $i"; break; fi
  i=$n
done
