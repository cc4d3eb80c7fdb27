// stand-in: /root/reference/include/SuperPoint.h includes it, src/SuperPoint.cc uses nothing of it.  TEST INFRASTRUCTURE.
#pragma once
