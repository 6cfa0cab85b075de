#!/bin/bash
cd "$(dirname "$0")/../.."
(cd integration/_build && timeout 100 ./channel_curv_demo 2 16 2; echo "rc=$?") 2>&1 | tail -5
(cd integration/_build && timeout 100 ./channel_curv_demo 2 32 2; echo "rc=$?") 2>&1 | tail -3
(cd integration/_build && timeout 100 ./channel_curv_demo 4 16 2; echo "rc=$?") 2>&1 | tail -3
