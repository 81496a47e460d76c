class File: pass
