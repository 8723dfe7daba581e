// TEST INFRASTRUCTURE: empty stand-in (include/stdafx.h includes it; nothing on the path uses Boost)
