"""Per-source-line profile of one kernel from an ncu report.

    cuobjdump -xelf all agh_b200/libagh_b200.so && nvdisasm -g -c decode.sm_100a.cubin > lines.txt
    ncu -i report.ncu-rep --page source --csv > src.csv
    python tools/ncu_lineprof.py <mangled kernel name> src.csv agh_b200/csrc/decode.cu lines.txt

Joins the SASS addresses of the report with nvdisasm's line table and prints, per source line, the share of stall
samples and of executed instructions."""
import re, csv, collections, sys
fn=sys.argv[1]; sass_csv=sys.argv[2]; srcfile=sys.argv[3]
txt=open(sys.argv[4] if len(sys.argv) > 4 else '/tmp/dec_all_lines.txt').read().split('\n')
start=None
for i,l in enumerate(txt):
    if l.startswith('.text.'+fn+':'): start=i; break
addr2line={}; cur=None
for l in txt[start+1:]:
    if l.strip().startswith('.section'): break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)),m.group(3).strip()); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m and cur: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(sass_csv)))
hdr=rows[1]
iS=hdr.index('# Samples'); iI=hdr.index('Instructions Executed')
data=[r for r in rows[2:] if len(r)==len(hdr) and r[iS].isdigit()]
base=int(data[0][0],16)
data=data[:len(addr2line)]
agg=collections.defaultdict(lambda:[0,0,0]); tot_s=tot_i=0
for r in data:
    a=int(r[0],16)-base
    key=addr2line.get(a,('?',0,''))
    k=key
    mm=re.findall(r'inlined at "([^"]+)", line (\d+)',key[2]) if key[2] else None
    if mm: k=(mm[-1][0].split('/')[-1],int(mm[-1][1]),'')
    kk=(k[0],k[1])
    agg[kk][0]+=int(r[iS]); agg[kk][1]+=int(r[iI]); agg[kk][2]+=1
    tot_s+=int(r[iS]); tot_i+=int(r[iI])
print('tot samples',tot_s,'inst',tot_i)
src=open(srcfile).read().split('\n')
for (f,ln),(s,i,c) in sorted(agg.items(), key=lambda x:(x[0][0],x[0][1])):
    if s/tot_s>0.003:
        print('%-12s %4d samp %.3f inst %.3f n=%3d | %s'%(f,ln,s/tot_s,i/tot_i,c,(src[ln-1].strip()[:90] if f==srcfile.split('/')[-1] and ln<=len(src) else '')))
