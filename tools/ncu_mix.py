"""Instruction mix and stall-sample distribution of one kernel from an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_mix.py rep"""
import csv, collections, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; data=rows[2:]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSm=hdr.index('# Samples')
ops=collections.Counter(); samp=collections.Counter(); tot=0; ts=0
for r in data:
    try: e=int(r[iE]); s=int(r[iSm])
    except: continue
    parts=r[iS].strip().split()
    op=parts[1] if parts[0].startswith('@') else parts[0]
    op=op.split('.')[0]
    ops[op]+=e; samp[op]+=s; tot+=e; ts+=s
print('total instr',tot,'samples',ts)
for op,c in ops.most_common(18): print('%-8s %12d %5.1f%%  samples %6d %5.1f%%'%(op,c,c/tot*100,samp[op],samp[op]/ts*100))
