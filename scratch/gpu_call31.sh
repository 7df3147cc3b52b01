#!/bin/bash
python scratch/fn_host_time.py 2>&1 | tail -5
