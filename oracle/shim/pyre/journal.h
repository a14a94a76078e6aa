// TEST INFRASTRUCTURE (oracle/_ref build only).  Empty stand-in for pyre's
// journal header (pyre 1.12.5 is not installed); the reference sources compiled
// into oracle/_ref include it (core/DateTime.cpp:12) but never use a channel.
#pragma once
