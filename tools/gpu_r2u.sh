#!/usr/bin/env bash
timeout 300 python tools/host_overhead.py 2>&1 | tail -6
