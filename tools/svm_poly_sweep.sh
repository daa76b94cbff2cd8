#!/bin/bash
# sweep the share of exponentials the SVM epilogue computes on the FMA pipe (CDS_SVM_POLY of every 8)
for p in 0 1 2 3 4; do
  echo -n "poly $p: "
  CDS_SVM_POLY=$p timeout 120 python tools/kernel_bench.py --kernel svm 2>&1 | grep svm_rbf9 | sed -e 's/.*"ms_median": \([0-9.]*\).*/\1 ms/'
done
