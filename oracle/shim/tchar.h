/* empty stand-in: the reference's targetver.h includes this Windows header (via stdafx.h) */
