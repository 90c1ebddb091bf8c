#!/bin/bash
timeout 300 python scratch/tma_coresidency_repro.py eig 2>&1 | tail -6
GWBSE_NO_TMA=1 timeout 300 python scratch/tma_coresidency_repro.py eig 2>&1 | tail -4
