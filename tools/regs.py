import re,subprocess,sys
for f in sys.argv[1:]:
    txt=open(f).read()
    ents=re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n(?:.*?\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
    for name,stack,ss,sl,regs in ents:
        dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
        dem=re.sub(r'\(.*','',dem)
        print(f"{regs:>4} regs stack {stack:>5} spill {ss}/{sl}  {dem[:110]}")
