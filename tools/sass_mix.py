"""Static SASS instruction mix per out-of-line routine of one kernel of a built library (cuobjdump -sass, split at RET/EXIT):
  python tools/sass_mix.py pbc_b200/libpbc_b200.so [mangled kernel name]
Used to see where ptxas put moves / additions (IMAD.MOV, IMAD.X on the multiplier pipe) before spending GPU time."""
import sys,re,subprocess,collections
lib=sys.argv[1]
kernel=sys.argv[2] if len(sys.argv) > 2 else '_ZN7pbcb20012k_f_miller_sILi128EEEvPKhS2_PjS3_S3_mmPKjm'
out=subprocess.run(['cuobjdump','-sass','-fun',kernel,lib],capture_output=True,text=True).stdout
ins=[re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$','',re.sub(r'^\s+/\*[0-9a-f]+\*/\s+','',l)) for l in out.splitlines() if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/',l)]
def key(s):
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',s); op=m.group(2); p=op.split('.')
    if p[0]=='IMAD':
        for t in ('WIDE','MOV','IADD','HI','X'):
            if t in p: return 'IMAD.'+t
    return p[0]
seg=[];cur=[]
for s in ins:
    cur.append(s)
    if s.startswith('RET') or s.startswith('EXIT'): seg.append(cur);cur=[]
print(len(ins))
for i,sg in enumerate(seg):
    if len(sg)<250: continue
    c=collections.Counter(key(s) for s in sg)
    print(i,len(sg),' '.join(f"{a}:{b}" for a,b in c.most_common(9)))
